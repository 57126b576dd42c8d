/* TEST INFRASTRUCTURE -- CPU oracle (see vlo_oracle.h; PARITY UNPINNED: algorithm lives in the
 * un-vendored `loam` dependency, gtsam_fusion/package.xml:26, README.md:21-25).
 *
 * Restates the map side of BasicLaserMapping::process of the LOAM fork
 * with the reference's knobs cornerFilterSize / surfaceFilterSize / mapCubeSize / mapDimensionsInCubes /
 * mapStartLocationInCubes / numNeighborSubmapCubes (gtsam_fusion/config/carla/loam_params.yaml:47-52):
 *   1. centre cube of transformTobeMapped, window shift while it is within 3 cubes of an edge
 *   2. neighbourhood cubes (+-numNeighborSubmapCubes) that pass the field-of-view test -> sub-map clouds
 *   3. VoxelGrid down-sampling of the corner / surface stacks (0.2 / 0.4 m)
 *   4. optimizeTransformTobeMapped against the sub-map (laser_mapping.c), needs > 10 corner and > 100 surface points
 *   5. the down-sampled stacks, moved with the optimised pose, are inserted into their cubes and the
 *      cubes are voxel-filtered again (corner 0.2 / surface 0.4 m)
 *
 * Frozen where upstream leaves it to PCL / container order (DESIGN.md "Choices frozen by the oracle"):
 *   M1  a map voxel = (cube, voxel coordinates floor(p / leaf)) -- per-cube VoxelGrid as upstream, so a voxel that
 *       straddles a cube boundary is filtered separately on either side; the filter is applied to every touched
 *       voxel right at insertion (upstream: only to the cubes of the FOV-valid neighbourhood, once per tick --
 *       identical whenever the touched cube is in that neighbourhood)
 *   M2  the stack is down-sampled in the sensor frame directly (upstream moves it to the map frame and back
 *       first, a float round trip)
 *   M3  re-filtering a voxel that already holds a centroid c with k new points gives (c + sum p) / (1 + k), the
 *       sum taken over offsets from the voxel origin quantised to 2^-20 m (order-free, V2 of scan_registration.c)
 *   M4  map point index (the k-NN tie-break) = order of voxel creation; the sub-map lists points by ascending index
 *   M5  a voxel keeps the identity (cube, voxel) it was created with; cubes shifted out of the window are dropped
 */
#include "vlo_oracle.h"
#include "detmath.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define LM_Q 1048576.0f
#define LM_EMPTY 0xFFFFFFFFFFFFFFFFull
#define LM_DEAD (1 << 30)
#define LM_VOX_BIAS (1 << 18)

void orc_mapping_register(const orc_config *c,
                          const orc_pt *corner_q, int n_cq, const orc_pt *surf_q, int n_sq,
                          const orc_pt *corner_map, int n_cm, const orc_pt *surf_map, int n_sm,
                          const float *seed, int use_kdtree,
                          orc_reg_result *res, int *trace_idx, float *trace_T);
void orc_point_to_map(const float *T, const orc_pt *in, int n, orc_pt *out);

struct orc_lmap {
    orc_config cfg;
    int cap;
    int cen[3];                 /* laserCloudCenWidth / Height / Depth (moves with window shifts) */
    int n[2];
    orc_pt *pts[2];             /* by id: centroid, w = 0 */
    int *cube[2];               /* by id: packed absolute cube coordinates | LM_DEAD */
    unsigned long long *keys[2]; int *vals[2]; int ts;
    /* scratch of the last process call */
    int *sub_ids[2]; int n_sub[2];
};

/* upstream: int((x + size/2) / size) [+ cen]; if (x + size/2 < 0) --  */
static int cube_coord(float x, float half, float size)
{
    float v = x + half;
    int c = (int)(v / size);
    if (v < 0.0f) c--;
    return c;
}

static int pack_cube(int ci, int cj, int ck) { return (ci + 512) | ((cj + 512) << 10) | ((ck + 512) << 20); }
static void unpack_cube(int v, int *c) { c[0] = (v & 1023) - 512; c[1] = ((v >> 10) & 1023) - 512; c[2] = ((v >> 20) & 1023) - 512; }

/* key of a map voxel: voxel coordinates (19 bit each, biased) + the offset of the point's cube from the cube of the
 * voxel origin (2 bit each): (voxel, cube) identity in 63 bits.  returns 0 if out of range */
static int voxel_key(const float *p, float leaf, float inv, float half, float size, const int *cube, int *vox, unsigned long long *key)
{
    unsigned long long k = 0;
    for (int a = 0; a < 3; a++) {
        int iv = (int)floorf(p[a] * inv);
        if (iv < -LM_VOX_BIAS || iv >= LM_VOX_BIAS) return 0;
        vox[a] = iv;
        int co = cube_coord((float)iv * leaf, half, size);
        k = (k << 21) | ((unsigned long long)(unsigned)(iv + LM_VOX_BIAS) << 2) | (unsigned long long)((cube[a] - co) & 3);
    }
    *key = k;
    return 1;
}

orc_lmap *orc_lmap_create(const orc_config *c, int cap)
{
    orc_lmap *m = (orc_lmap *)calloc(1, sizeof(orc_lmap));
    m->cfg = *c; m->cap = cap;
    for (int a = 0; a < 3; a++) m->cen[a] = c->map_start_cubes[a];
    m->ts = 1; while (m->ts < 2 * cap) m->ts <<= 1;
    for (int w = 0; w < 2; w++) {
        m->pts[w] = (orc_pt *)calloc((size_t)cap, sizeof(orc_pt));
        m->cube[w] = (int *)calloc((size_t)cap, sizeof(int));
        m->keys[w] = (unsigned long long *)malloc(sizeof(unsigned long long) * (size_t)m->ts);
        m->vals[w] = (int *)malloc(sizeof(int) * (size_t)m->ts);
        for (int i = 0; i < m->ts; i++) { m->keys[w][i] = LM_EMPTY; m->vals[w][i] = -1; }
        m->sub_ids[w] = (int *)malloc(sizeof(int) * (size_t)cap);
    }
    return m;
}

void orc_lmap_free(orc_lmap *m)
{
    if (!m) return;
    for (int w = 0; w < 2; w++) { free(m->pts[w]); free(m->cube[w]); free(m->keys[w]); free(m->vals[w]); free(m->sub_ids[w]); }
    free(m);
}

int orc_lmap_size(const orc_lmap *m, int which) { return m->n[which]; }
void orc_lmap_get(const orc_lmap *m, int which, orc_pt *pts, int *cube)
{
    if (pts) memcpy(pts, m->pts[which], sizeof(orc_pt) * (size_t)m->n[which]);
    if (cube) memcpy(cube, m->cube[which], sizeof(int) * (size_t)m->n[which]);
}
void orc_lmap_window(const orc_lmap *m, int *cen3) { for (int a = 0; a < 3; a++) cen3[a] = m->cen[a]; }
int orc_lmap_submap(const orc_lmap *m, int which, int *ids)
{
    if (ids) memcpy(ids, m->sub_ids[which], sizeof(int) * (size_t)m->n_sub[which]);
    return m->n_sub[which];
}

static unsigned long long mix64(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

/* insertion step 5 for one cloud: pts (sensor frame) moved with T */
static void lmap_insert(orc_lmap *m, int w, const orc_pt *pts, int n, const float *T)
{
    const orc_config *c = &m->cfg;
    const float leaf = w == 0 ? c->corner_filter_size : c->surface_filter_size, inv = 1.0f / leaf;
    const float size = c->map_cube_size, half = size / 2.0f;
    if (n <= 0) return;
    orc_pt *pm = (orc_pt *)malloc(sizeof(orc_pt) * (size_t)n);
    orc_point_to_map(T, pts, n, pm);
    int *touched = (int *)malloc(sizeof(int) * (size_t)n);   /* ids in order of first touch */
    int n_touched = 0;
    typedef struct { long long s[3]; int k; int fresh; int vox[3]; } acc_t;
    acc_t *acc = (acc_t *)calloc((size_t)n, sizeof(acc_t));
    int *slot_acc = (int *)malloc(sizeof(int) * (size_t)m->cap);   /* id -> index into acc, -1 */
    /* only ids touched this call are looked at: initialise lazily through a stamp array */
    memset(slot_acc, 0xFF, sizeof(int) * (size_t)m->cap);
    for (int q = 0; q < n; q++) {
        float p[3] = { pm[q].x, pm[q].y, pm[q].z };
        int cube[3], vox[3], inwin = 1;
        for (int a = 0; a < 3; a++) {
            cube[a] = cube_coord(p[a], half, size);
            int rel = cube[a] + m->cen[a];
            if (rel < 0 || rel >= c->map_dims[a] || cube[a] < -512 || cube[a] > 511) inwin = 0;
        }
        if (!inwin) continue;
        unsigned long long key;
        if (!voxel_key(p, leaf, inv, half, size, cube, vox, &key)) continue;
        int slot = (int)(mix64(key) & (unsigned long long)(m->ts - 1));
        while (m->keys[w][slot] != LM_EMPTY && m->keys[w][slot] != key) slot = (slot + 1) & (m->ts - 1);
        int id = -1, fresh = 0;
        if (m->keys[w][slot] == key) id = m->vals[w][slot];
        if (id < 0) {
            if (m->n[w] >= m->cap) continue;                 /* capacity: dropped */
            id = m->n[w]++;
            m->keys[w][slot] = key; m->vals[w][slot] = id;
            m->cube[w][id] = pack_cube(cube[0], cube[1], cube[2]);
            fresh = 1;
        } else if (slot_acc[id] < 0 && (m->cube[w][id] & LM_DEAD)) {
            m->cube[w][id] &= ~LM_DEAD;                      /* voxel of an evicted cube comes back empty */
            fresh = 1;
        }
        int ai = slot_acc[id];
        if (ai < 0) {
            ai = n_touched; slot_acc[id] = ai; touched[n_touched++] = id;
            acc[ai].fresh = fresh;
            for (int a = 0; a < 3; a++) acc[ai].vox[a] = vox[a];
        }
        for (int a = 0; a < 3; a++) {
            float o = (float)vox[a] * leaf;
            acc[ai].s[a] += (long long)(int)rintf((p[a] - o) * LM_Q);
        }
        acc[ai].k++;
    }
    for (int t = 0; t < n_touched; t++) {
        int id = touched[t];
        acc_t *A = &acc[t];
        float old[3] = { m->pts[w][id].x, m->pts[w][id].y, m->pts[w][id].z }, out[3];
        int cnt = A->k + (A->fresh ? 0 : 1);
        for (int a = 0; a < 3; a++) {
            float o = (float)A->vox[a] * leaf;
            long long s = A->s[a];
            if (!A->fresh) s += (long long)(int)rintf((old[a] - o) * LM_Q);
            out[a] = o + ((float)(int)s / (float)cnt) * (1.0f / 1048576.0f);
        }
        m->pts[w][id].x = out[0]; m->pts[w][id].y = out[1]; m->pts[w][id].z = out[2]; m->pts[w][id].w = 0.0f;
    }
    free(pm); free(touched); free(acc); free(slot_acc);
}

/* the upstream insertion step on its own (e.g. to preload a prior map) */
void orc_lmap_insert(orc_lmap *m, const orc_pt *corner, int nc, const orc_pt *surf, int ns, const float *T)
{
    lmap_insert(m, 0, corner, nc, T);
    lmap_insert(m, 1, surf, ns, T);
}

/* steps 1-2: window shift around T, eviction, FOV-valid neighbourhood mask ((2 nb + 1)^3 bytes, x fastest),
 * centre cube (absolute coordinates) */
void orc_lmap_select(orc_lmap *m, const float *T, int *centre_abs, uint8_t *mask)
{
    const orc_config *c = &m->cfg;
    const float size = c->map_cube_size, half = size / 2.0f;
    const int nb = c->n_neighbor_cubes, side = 2 * nb + 1;
    int shifted = 0;
    for (int a = 0; a < 3; a++) {
        int ca = cube_coord(T[3 + a], half, size);
        int cc = ca + m->cen[a];
        while (cc < 3) { cc++; m->cen[a]++; shifted = 1; }
        while (cc >= c->map_dims[a] - 3) { cc--; m->cen[a]--; shifted = 1; }
        centre_abs[a] = ca;
    }
    if (shifted) {
        for (int w = 0; w < 2; w++)
            for (int id = 0; id < m->n[w]; id++) {
                int cb[3]; unpack_cube(m->cube[w][id], cb);
                for (int a = 0; a < 3; a++) { int rel = cb[a] + m->cen[a]; if (rel < 0 || rel >= c->map_dims[a]) m->cube[w][id] |= LM_DEAD; }
            }
    }
    /* pointOnYAxis = pointAssociateToMap((0, 10, 0)) */
    orc_pt yin = { 0.0f, 10.0f, 0.0f, 0.0f }, yax;
    orc_point_to_map(T, &yin, 1, &yax);
    const float s3 = 10.0f * sqrtf(3.0f);
    for (int dk = 0; dk < side; dk++) for (int dj = 0; dj < side; dj++) for (int di = 0; di < side; di++) {
        int cb[3] = { centre_abs[0] - nb + di, centre_abs[1] - nb + dj, centre_abs[2] - nb + dk };
        int ok = 1;
        for (int a = 0; a < 3; a++) { int rel = cb[a] + m->cen[a]; if (rel < 0 || rel >= c->map_dims[a]) ok = 0; }
        int fov = 0;
        if (ok) {
            float cx = size * (float)cb[0], cy = size * (float)cb[1], cz = size * (float)cb[2];
            for (int ii = -1; ii <= 1 && !fov; ii += 2) for (int jj = -1; jj <= 1 && !fov; jj += 2) for (int kk = -1; kk <= 1 && !fov; kk += 2) {
                float px = cx + half * (float)ii, py = cy + half * (float)jj, pz = cz + half * (float)kk;
                float ax = T[3] - px, ay = T[4] - py, az = T[5] - pz;
                float bx = yax.x - px, by = yax.y - py, bz = yax.z - pz;
                float s1 = (ax * ax + ay * ay) + az * az;
                float s2 = (bx * bx + by * by) + bz * bz;
                float r = s3 * sqrtf(s1);
                float check1 = ((100.0f + s1) - s2) - r;
                float check2 = ((100.0f + s1) - s2) + r;
                if (check1 < 0.0f && check2 > 0.0f) fov = 1;
            }
        }
        mask[(dk * side + dj) * side + di] = (uint8_t)fov;
    }
    /* sub-map = live points of the masked cubes, ascending id */
    for (int w = 0; w < 2; w++) {
        int ns = 0;
        for (int id = 0; id < m->n[w]; id++) {
            int v = m->cube[w][id];
            if (v & LM_DEAD) continue;
            int cb[3]; unpack_cube(v, cb);
            int r0 = cb[0] - (centre_abs[0] - nb), r1 = cb[1] - (centre_abs[1] - nb), r2 = cb[2] - (centre_abs[2] - nb);
            if (r0 < 0 || r0 >= side || r1 < 0 || r1 >= side || r2 < 0 || r2 >= side) continue;
            if (mask[(r2 * side + r1) * side + r0]) m->sub_ids[w][ns++] = id;
        }
        m->n_sub[w] = ns;
    }
}

/* one BasicLaserMapping::process().  corner/surf stack: the sweep's less-sharp / less-flat clouds (sensor frame at
 * sweep end); seed: transformTobeMapped after transformAssociateToMap.  info[6] = n_ds corner, n_ds surf,
 * n_sub corner, n_sub surf, n_map corner, n_map surf (after insertion). */
void orc_lmap_process(orc_lmap *m, const orc_pt *corner_stack, int nc, const orc_pt *surf_stack, int ns,
                      const float *seed, orc_reg_result *res, int *info)
{
    const orc_config *c = &m->cfg;
    const int nb = c->n_neighbor_cubes, side = 2 * nb + 1;
    int centre[3];
    uint8_t *mask = (uint8_t *)malloc((size_t)side * side * side);
    orc_lmap_select(m, seed, centre, mask);
    orc_pt *cds = (orc_pt *)malloc(sizeof(orc_pt) * (size_t)(nc > 0 ? nc : 1));
    orc_pt *sds = (orc_pt *)malloc(sizeof(orc_pt) * (size_t)(ns > 0 ? ns : 1));
    int ncd = orc_voxel_downsample(corner_stack, nc, c->corner_filter_size, cds);
    int nsd = orc_voxel_downsample(surf_stack, ns, c->surface_filter_size, sds);
    orc_pt *subc = (orc_pt *)malloc(sizeof(orc_pt) * (size_t)(m->n_sub[0] > 0 ? m->n_sub[0] : 1));
    orc_pt *subs = (orc_pt *)malloc(sizeof(orc_pt) * (size_t)(m->n_sub[1] > 0 ? m->n_sub[1] : 1));
    for (int i = 0; i < m->n_sub[0]; i++) subc[i] = m->pts[0][m->sub_ids[0][i]];
    for (int i = 0; i < m->n_sub[1]; i++) subs[i] = m->pts[1][m->sub_ids[1][i]];
    orc_mapping_register(c, cds, ncd, sds, nsd, subc, m->n_sub[0], subs, m->n_sub[1], seed, 1, res, NULL, NULL);
    lmap_insert(m, 0, cds, ncd, res->transform);
    lmap_insert(m, 1, sds, nsd, res->transform);
    if (info) { info[0] = ncd; info[1] = nsd; info[2] = m->n_sub[0]; info[3] = m->n_sub[1]; info[4] = m->n[0]; info[5] = m->n[1]; }
    free(mask); free(cds); free(sds); free(subc); free(subs);
}
