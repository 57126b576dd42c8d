/* TEST INFRASTRUCTURE -- CPU oracle (see vlo_oracle.h header comment; parity unpinned).
 *
 * Restates MultiScanRegistration::process + BasicScanRegistration::extractFeatures of the LOAM
 * fork the reference depends on (knobs from
 * gtsam_fusion/config/carla/loam_params.yaml:3,22-31).  Sequential, one ring / one sector at a
 * time, exactly in upstream's evaluation order: this is the spec the CUDA kernels are held to.
 */
#include "vlo_oracle.h"
#include "detmath.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_PI 3.14159265358979323846

void orc_default_config(orc_config *c)
{
    c->scan_period = 0.1f;
    c->n_rings = 16; c->lower_deg = -15.0f; c->upper_deg = 15.0f;
    c->feature_regions = 6;
    c->curvature_region = 5;
    c->max_corner_sharp = 2;
    c->max_corner_less_sharp = 20;
    c->max_surface_flat = 4;
    c->surface_curvature_threshold = 0.1f;
    c->less_flat_filter_size = 0.2f;
    c->odom_max_iterations = 25;
    c->odom_delta_t_abort = 0.05f;
    c->odom_delta_r_abort = 0.05f;
    c->odom_degen_eig = 30.0f;
    c->map_max_iterations = 10;
    c->map_delta_t_abort = 0.05f;
    c->map_delta_r_abort = 0.05f;
    c->map_degen_eig = 40.0f;
    c->deskew = 1;
    c->odom_forward_bound_quirk = 0;
    c->dopt_rot_threshold = 11.5f;
    c->dopt_trans_threshold = 28.9f;
    c->corner_filter_size = 0.2f;
    c->surface_filter_size = 0.4f;
    c->map_cube_size = 10.0f;
    c->map_dims[0] = 101; c->map_dims[1] = 51; c->map_dims[2] = 101;
    c->map_start_cubes[0] = 50; c->map_start_cubes[1] = 25; c->map_start_cubes[2] = 50;
    c->n_neighbor_cubes = 5;
    c->io_ratio = 2;
    c->rotate_input = 0; c->input_rotation[0] = c->input_rotation[1] = c->input_rotation[2] = 0.0f;
    c->ring_field = -1; c->ring_field_type = 0;
}

/* MultiScanMapper::getRingForAngle: int(((angle*180/M_PI) - lower) * factor + 0.5) */
static int ring_for_angle(const orc_config *c, float angle)
{
    float factor = (float)(c->n_rings - 1) / (c->upper_deg - c->lower_deg);
    float a180 = angle * 180.0f;
    double v = ((double)a180 / ORC_PI - (double)c->lower_deg) * (double)factor + 0.5;
    return (int)v;
}

/* rotateInputCloud / inputCloudRotation (gtsam_fusion/config/.../loam_params.yaml:4-5; the code is in the un-vendored fork):
 * frozen here as R = Rz(yaw) Ry(pitch) Rx(roll) applied to the ROS-frame point, matrix entries in double then rounded to
 * float, products summed as (r0 x + r1 y) + r2 z. */
static void input_rotation_matrix(const orc_config *c, float R[9])
{
    double cy = cos((double)c->input_rotation[0]), sy = sin((double)c->input_rotation[0]);
    double cp = cos((double)c->input_rotation[1]), sp = sin((double)c->input_rotation[1]);
    double cr = cos((double)c->input_rotation[2]), sr = sin((double)c->input_rotation[2]);
    R[0] = (float)(cy * cp); R[1] = (float)(cy * sp * sr - sy * cr); R[2] = (float)(cy * sp * cr + sy * sr);
    R[3] = (float)(sy * cp); R[4] = (float)(sy * sp * sr + cy * cr); R[5] = (float)(sy * sp * cr - cy * sr);
    R[6] = (float)(-sp);     R[7] = (float)(cp * sr);                R[8] = (float)(cp * cr);
}
/* ROS-frame x y z of a raw point, rotated when rotateInputCloud is set */
static inline void ros_point(const orc_config *c, const float *R, const float *p, float *rx, float *ry, float *rz)
{
    float x = p[0], y = p[1], z = p[2];
    if (c->rotate_input) {
        float ax = (R[0] * x + R[1] * y) + R[2] * z;
        float ay = (R[3] * x + R[4] * y) + R[5] * z;
        float az = (R[6] * x + R[7] * y) + R[8] * z;
        x = ax; y = ay; z = az;
    }
    *rx = x; *ry = y; *rz = z;
}

int orc_organise(const orc_config *c, const float *raw, int n, int stride,
                 orc_pt *out, int *ring_start, int *src_index)
{
    float Rm[9];
    input_rotation_matrix(c, Rm);
    const int R = c->n_rings;
    for (int r = 0; r <= R; r++) ring_start[r] = 0;
    if (n <= 0) return 0;
    int *ring = (int *)malloc(sizeof(int) * (size_t)n);
    float *inten = (float *)malloc(sizeof(float) * (size_t)n);

    float fx, fy, fz, lx, ly, lz;
    ros_point(c, Rm, raw, &fx, &fy, &fz);
    ros_point(c, Rm, raw + (size_t)(n - 1) * stride, &lx, &ly, &lz);
    float startOri = -orc_atan2f(fy, fx);
    float endOri = -orc_atan2f(ly, lx) + 2.0f * (float)ORC_PI;
    if ((double)(endOri - startOri) > 3 * ORC_PI) endOri = (float)((double)endOri - 2 * ORC_PI);
    else if ((double)(endOri - startOri) < ORC_PI) endOri = (float)((double)endOri + 2 * ORC_PI);

    int halfPassed = 0;
    int *count = (int *)calloc((size_t)R, sizeof(int));
    for (int i = 0; i < n; i++) {
        const float *p = raw + (size_t)i * stride;
        float rx, ry, rz;
        ros_point(c, Rm, p, &rx, &ry, &rz);
        float x = ry, y = rz, z = rx;                /* LOAM frame: x<-y, y<-z, z<-x */
        ring[i] = -1;
        if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;
        if ((x * x + y * y) + z * z < 0.0001f) continue;
        int id;
        if (c->ring_field >= 0) {                    /* ring id as delivered by the driver (useCloudIntensityandRingFields) */
            float rf;
            if (c->ring_field_type == 0) rf = p[c->ring_field];
            else {
                const unsigned char *bp = (const unsigned char *)p + c->ring_field;
                rf = c->ring_field_type == 1 ? (float)(bp[0] | (bp[1] << 8)) : (float)bp[0];
            }
            if (!(rf >= 0.0f && rf < (float)R)) continue;       /* also rejects NaN */
            id = (int)rf;
        } else {
            float angle = orc_atanf(y / sqrtf(x * x + z * z));
            id = ring_for_angle(c, angle);
        }
        if (id >= R || id < 0) continue;
        float ori = -orc_atan2f(x, z);
        if (!halfPassed) {
            if ((double)ori < (double)startOri - ORC_PI / 2) ori = (float)((double)ori + 2 * ORC_PI);
            else if ((double)ori > (double)startOri + ORC_PI * 3 / 2) ori = (float)((double)ori - 2 * ORC_PI);
            if ((double)(ori - startOri) > ORC_PI) halfPassed = 1;
        } else {
            ori = (float)((double)ori + 2 * ORC_PI);
            if ((double)ori < (double)endOri - ORC_PI * 3 / 2) ori = (float)((double)ori + 2 * ORC_PI);
            else if ((double)ori > (double)endOri + ORC_PI / 2) ori = (float)((double)ori - 2 * ORC_PI);
        }
        float relTime = c->scan_period * (ori - startOri) / (endOri - startOri);
        ring[i] = id;
        inten[i] = (float)id + relTime;
        count[id]++;
    }
    for (int r = 0; r < R; r++) ring_start[r + 1] = ring_start[r] + count[r];
    int *cursor = (int *)malloc(sizeof(int) * (size_t)R);
    memcpy(cursor, ring_start, sizeof(int) * (size_t)R);
    for (int i = 0; i < n; i++) {
        if (ring[i] < 0) continue;
        const float *p = raw + (size_t)i * stride;
        int o = cursor[ring[i]]++;
        float rx, ry, rz;
        ros_point(c, Rm, p, &rx, &ry, &rz);
        out[o].x = ry; out[o].y = rz; out[o].z = rx; out[o].w = inten[i];
        if (src_index) src_index[o] = i;
    }
    int total = ring_start[R];
    free(ring); free(inten); free(count); free(cursor);
    return total;
}

static inline float sqdiff(const orc_pt *a, const orc_pt *b)
{
    float dx = a->x - b->x, dy = a->y - b->y, dz = a->z - b->z;
    return (dx * dx + dy * dy) + dz * dz;
}
static inline float sqdiff_w(const orc_pt *a, const orc_pt *b, float wb)
{
    float dx = a->x - b->x * wb, dy = a->y - b->y * wb, dz = a->z - b->z * wb;
    return (dx * dx + dy * dy) + dz * dz;
}
static inline float sqnorm(const orc_pt *a) { return (a->x * a->x + a->y * a->y) + a->z * a->z; }

/* BasicScanRegistration::markAsPicked */
static void mark_as_picked(const orc_pt *cloud, uint8_t *picked, int K, int idx, int scan_start)
{
    int s = idx - scan_start;
    picked[s] = 1;
    for (int i = 1; i <= K; i++) {
        if (sqdiff(&cloud[idx + i], &cloud[idx + i - 1]) > 0.05) break;   /* double literal upstream */
        picked[s + i] = 1;
    }
    for (int i = 1; i <= K; i++) {
        if (sqdiff(&cloud[idx - i], &cloud[idx - i + 1]) > 0.05) break;
        picked[s - i] = 1;
    }
}

/* pcl::VoxelGrid<PointXYZI> (leaf = lessFlatFilterSize, loam_params.yaml:31) restated: points are
 * grouped by voxel (floor(p * 1/leaf) per axis) and each voxel emits the centroid of its members
 * (all four fields).  Two things PCL leaves to std::sort / float summation order are FROZEN here so
 * that a parallel implementation can be bit-exact:
 *   V1  output order = order of first appearance of the voxel along the ring (PCL: ascending linear
 *       voxel index, an artefact of its sort key);
 *   V2  centroid = voxel origin + mean of the members' offsets, the offsets quantised to 2^-20 m and
 *       summed as integers (order independent; <= 0.5 um from any float-order sum), the mean taken
 *       as one float32 division (float)sum / (float)count;
 *       intensity likewise on its fractional part (rel-time), integer part = ring. */
typedef struct { int ix, iy, iz; int first; int cnt; long long sx, sy, sz, sw; } vox_grp;
#define VOX_Q 1048576.0f
static int voxel_downsample(const orc_pt *in, int n, float leaf, orc_pt *out)
{
    if (n == 0) return 0;
    float inv = 1.0f / leaf;
    int cap = 1; while (cap < 2 * n) cap <<= 1;
    int *table = (int *)malloc(sizeof(int) * (size_t)cap);
    for (int i = 0; i < cap; i++) table[i] = -1;
    vox_grp *g = (vox_grp *)malloc(sizeof(vox_grp) * (size_t)n);
    int m = 0;
    for (int i = 0; i < n; i++) {
        int ix = (int)floorf(in[i].x * inv), iy = (int)floorf(in[i].y * inv), iz = (int)floorf(in[i].z * inv);
        unsigned h = ((unsigned)ix * 73856093u) ^ ((unsigned)iy * 19349663u) ^ ((unsigned)iz * 83492791u);
        int slot = (int)(h & (unsigned)(cap - 1));
        int gi = -1;
        while (table[slot] >= 0) {
            vox_grp *c = &g[table[slot]];
            if (c->ix == ix && c->iy == iy && c->iz == iz) { gi = table[slot]; break; }
            slot = (slot + 1) & (cap - 1);
        }
        if (gi < 0) {
            gi = m++;
            table[slot] = gi;
            g[gi].ix = ix; g[gi].iy = iy; g[gi].iz = iz; g[gi].first = i; g[gi].cnt = 0;
            g[gi].sx = g[gi].sy = g[gi].sz = g[gi].sw = 0;
        }
        float ox = (float)ix * leaf, oy = (float)iy * leaf, oz = (float)iz * leaf;
        g[gi].sx += (long long)(int)rintf((in[i].x - ox) * VOX_Q);
        g[gi].sy += (long long)(int)rintf((in[i].y - oy) * VOX_Q);
        g[gi].sz += (long long)(int)rintf((in[i].z - oz) * VOX_Q);
        g[gi].sw += (long long)(int)rintf((in[i].w - (float)(int)in[i].w) * VOX_Q);
        g[gi].cnt++;
    }
    for (int k = 0; k < m; k++) {
        float c = (float)g[k].cnt;
        const float q = 1.0f / 1048576.0f;
        float ox = (float)g[k].ix * leaf, oy = (float)g[k].iy * leaf, oz = (float)g[k].iz * leaf;
        out[k].x = ox + ((float)(int)g[k].sx / c) * q;
        out[k].y = oy + ((float)(int)g[k].sy / c) * q;
        out[k].z = oz + ((float)(int)g[k].sz / c) * q;
        out[k].w = (float)(int)in[g[k].first].w + ((float)(int)g[k].sw / c) * q;
    }
    free(table); free(g);
    return m;
}

/* leaf <= 0: no filtering (copy) */
int orc_voxel_downsample(const orc_pt *in, int n, float leaf, orc_pt *out)
{
    if (!(leaf > 0.0f)) { memcpy(out, in, sizeof(orc_pt) * (size_t)n); return n; }
    return voxel_downsample(in, n, leaf, out);
}

void orc_extract(const orc_config *c, const orc_pt *cloud, const int *ring_start,
                 int8_t *label, float *curvature, uint8_t *picked_out,
                 int *sharp_idx, int *less_sharp_idx, int *flat_idx,
                 orc_pt *less_flat, int *lsharp_ring_start, int *lflat_ring_start,
                 orc_feature_counts *counts)
{
    const int R = c->n_rings, K = c->curvature_region, NR = c->feature_regions;
    const float thr = c->surface_curvature_threshold;
    int total = ring_start[R];
    memset(label, 0, (size_t)total);
    memset(curvature, 0, sizeof(float) * (size_t)total);
    if (picked_out) memset(picked_out, 0, (size_t)total);
    int n_sharp = 0, n_ls = 0, n_flat = 0, n_lf = 0;
    int max_ring = 0;
    for (int r = 0; r < R; r++) { int s = ring_start[r + 1] - ring_start[r]; if (s > max_ring) max_ring = s; }
    uint8_t *picked = (uint8_t *)malloc((size_t)max_ring + 16);
    int *sortidx = (int *)malloc(sizeof(int) * ((size_t)max_ring + 1));
    orc_pt *cand = (orc_pt *)malloc(sizeof(orc_pt) * ((size_t)max_ring + 1));

    for (int r = 0; r < R; r++) {
        lsharp_ring_start[r] = n_ls;
        lflat_ring_start[r] = n_lf;
        int start = ring_start[r], end = ring_start[r + 1] - 1;   /* inclusive, as _scanIndices */
        if (ring_start[r + 1] - ring_start[r] <= 0) continue;
        if (end <= start + 2 * K) continue;                       /* "skip empty scans" */
        int scan_size = end - start + 1;
        memset(picked, 0, (size_t)scan_size);
        int n_cand = 0;

        /* setScanBuffersFor: occlusion / parallel-beam rejection */
        for (int i = start + K; i < end - K; i++) {
            const orc_pt *prev = &cloud[i - 1], *pt = &cloud[i], *next = &cloud[i + 1];
            float diffNext = sqdiff(next, pt);
            if (diffNext > 0.1) {                 /* double literals upstream: compare in double */
                float depth1 = sqrtf(sqnorm(pt));
                float depth2 = sqrtf(sqnorm(next));
                if (depth1 > depth2) {
                    float wd = sqrtf(sqdiff_w(next, pt, depth2 / depth1)) / depth2;
                    if (wd < 0.1) {
                        for (int m = 0; m <= K; m++) picked[i - start - K + m] = 1;
                        continue;
                    }
                } else {
                    float wd = sqrtf(sqdiff_w(pt, next, depth1 / depth2)) / depth1;
                    if (wd < 0.1) {
                        for (int m = 0; m <= K; m++) picked[i - start + 1 + m] = 1;
                    }
                }
            }
            float diffPrev = sqdiff(pt, prev);
            float dis = sqnorm(pt);
            if (diffNext > 0.0002 * dis && diffPrev > 0.0002 * dis) picked[i - start] = 1;
        }

        for (int j = 0; j < NR; j++) {
            int sp = ((start + K) * (NR - j) + (end - K) * j) / NR;
            int ep = ((start + K) * (NR - 1 - j) + (end - K) * (j + 1)) / NR - 1;
            if (ep <= sp) continue;
            int region = ep - sp + 1;

            /* setRegionBuffersFor: curvature + stable insertion sort ascending */
            float w = (float)(-2 * K);
            for (int i = sp; i <= ep; i++) {
                float dx = w * cloud[i].x, dy = w * cloud[i].y, dz = w * cloud[i].z;
                for (int m = 1; m <= K; m++) {
                    dx += cloud[i + m].x + cloud[i - m].x;
                    dy += cloud[i + m].y + cloud[i - m].y;
                    dz += cloud[i + m].z + cloud[i - m].z;
                }
                curvature[i] = (dx * dx + dy * dy) + dz * dz;
                sortidx[i - sp] = i;
            }
            for (int i = 1; i < region; i++) {
                int v = sortidx[i];
                float cv = curvature[v];
                int j2 = i;
                while (j2 >= 1 && cv < curvature[sortidx[j2 - 1]]) { sortidx[j2] = sortidx[j2 - 1]; j2--; }
                sortidx[j2] = v;
            }

            /* corners: from the largest curvature down */
            int largest = 0;
            for (int k = region; k > 0 && largest < c->max_corner_less_sharp;) {
                int idx = sortidx[--k];
                if (picked[idx - start] == 0 && curvature[idx] > thr) {
                    largest++;
                    if (largest <= c->max_corner_sharp) {
                        label[idx] = 2;
                        sharp_idx[n_sharp++] = idx;
                    } else {
                        label[idx] = 1;
                    }
                    less_sharp_idx[n_ls++] = idx;
                    mark_as_picked(cloud, picked, K, idx, start);
                }
            }
            /* flats: from the smallest curvature up */
            int smallest = 0;
            for (int k = 0; k < region && smallest < c->max_surface_flat; k++) {
                int idx = sortidx[k];
                if (picked[idx - start] == 0 && curvature[idx] < thr) {
                    smallest++;
                    label[idx] = -1;
                    flat_idx[n_flat++] = idx;
                    mark_as_picked(cloud, picked, K, idx, start);
                }
            }
            /* less flat candidates: everything not a corner */
            for (int i = sp; i <= ep; i++)
                if (label[i] <= 0) cand[n_cand++] = cloud[i];
        }
        n_lf += voxel_downsample(cand, n_cand, c->less_flat_filter_size, less_flat + n_lf);
        if (picked_out) memcpy(picked_out + start, picked, (size_t)scan_size);
    }
    lsharp_ring_start[R] = n_ls;
    lflat_ring_start[R] = n_lf;
    /* rings skipped above still need monotone offsets */
    counts->n_sharp = n_sharp; counts->n_less_sharp = n_ls; counts->n_flat = n_flat; counts->n_less_flat = n_lf;
    free(picked); free(sortidx); free(cand);
}
