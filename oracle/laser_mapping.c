/* TEST INFRASTRUCTURE -- CPU oracle (see vlo_oracle.h; PARITY UNPINNED: algorithm lives in the
 * un-vendored `loam` dependency, gtsam_fusion/package.xml:26, README.md:21-25).
 *
 * Restates BasicLaserMapping::optimizeTransformTobeMapped of the LOAM fork
 * with the reference's knobs mapMaxIterations / mapDeltaTAbort / mapDeltaRAbort / mapDegenEigVal
 * (gtsam_fusion/config/carla/loam_params.yaml:44-46,53): per iteration every feature point of the
 * current sweep is moved into the map frame, its 5 nearest map points are found, corners fit a line
 * through the principal direction of the 5 (3x3 covariance eigen-decomposition), surfaces fit a plane
 * (5x3 least squares), the point-to-line / point-to-plane rows are reduced to the 6x6 normal equations
 * (R1 order of laser_odometry.c) and solved exactly like the odometry step.
 *
 * Frozen where Eigen leaves it open: 3x3 symmetric eigen-solver = cyclic Jacobi, 6 sweeps, pairs
 * (0,1),(0,2),(1,2); 5x3 least squares = column-pivoted Householder QR with recomputed column norms.
 */
#include "vlo_oracle.h"
#include "detmath.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define NTERM 28
void orc_reduce_r1(const float *terms, int Q, int nterm, float *total);
void orc_gn_update(const float *total, int iter, float degen_thr, float dT_abort, float dR_abort,
                   float *T, orc_reg_result *res, int *converged, int map_order);
void orc_finish_result(const orc_config *c, const float *total, int n_corr, orc_reg_result *res);

/* pointAssociateToMap: rotateZXY(rz, rx, ry) then + t; trig = srx crx sry cry srz crz */
static inline void to_map(const float *T, const float *trig, const orc_pt *pi, orc_pt *po)
{
    float x = pi->x, y = pi->y, z = pi->z;
    float sx = trig[0], cx = trig[1], sy = trig[2], cy = trig[3], sz = trig[4], cz = trig[5];
    float x0 = x; x = cz * x0 - sz * y; y = sz * x0 + cz * y;
    float y0 = y; y = cx * y0 - sx * z; z = sx * y0 + cx * z;
    x0 = x;       x = cy * x0 + sy * z; z = cy * z - sy * x0;
    po->x = x + T[3]; po->y = y + T[4]; po->z = z + T[5]; po->w = pi->w;
}

/* cyclic Jacobi on a symmetric 3x3 (full storage); eval ascending, evec[k] = column k as vector */
void orc_eig3_jacobi(const float *Ain, float *eval, float *evec)
{
    float A[3][3], V[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { A[i][j] = Ain[i * 3 + j]; V[i][j] = (i == j) ? 1.0f : 0.0f; }
    static const int PQ[3][2] = { {0, 1}, {0, 2}, {1, 2} };
    for (int sweep = 0; sweep < 6; sweep++) {
        for (int m = 0; m < 3; m++) {
            int p = PQ[m][0], q = PQ[m][1];
            float apq = A[p][q];
            if (apq == 0.0f) continue;
            float theta = (A[q][q] - A[p][p]) / (2.0f * apq);
            float t = 1.0f / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
            if (theta < 0.0f) t = -t;
            float c = 1.0f / sqrtf(t * t + 1.0f), s = t * c;
            for (int i = 0; i < 3; i++) { float aip = A[i][p], aiq = A[i][q]; A[i][p] = c * aip - s * aiq; A[i][q] = s * aip + c * aiq; }
            for (int j = 0; j < 3; j++) { float apj = A[p][j], aqj = A[q][j]; A[p][j] = c * apj - s * aqj; A[q][j] = s * apj + c * aqj; }
            for (int i = 0; i < 3; i++) { float vip = V[i][p], viq = V[i][q]; V[i][p] = c * vip - s * viq; V[i][q] = s * vip + c * viq; }
            A[q][p] = A[p][q];
            int r = 3 - p - q;
            A[r][p] = A[p][r]; A[r][q] = A[q][r];
        }
    }
    int order[3] = { 0, 1, 2 };
    for (int i = 1; i < 3; i++) {
        int v = order[i], j = i;
        while (j >= 1 && A[v][v] < A[order[j - 1]][order[j - 1]]) { order[j] = order[j - 1]; j--; }
        order[j] = v;
    }
    for (int k = 0; k < 3; k++) {
        eval[k] = A[order[k]][order[k]];
        for (int i = 0; i < 3; i++) evec[k * 3 + i] = V[i][order[k]];
    }
}

/* least squares A(5x3) x = b via column-pivoted Householder QR */
void orc_lstsq53(const float *Ain, const float *bin, float *x)
{
    float A[5][3], b[5];
    int perm[3] = { 0, 1, 2 };
    for (int i = 0; i < 5; i++) { for (int j = 0; j < 3; j++) A[i][j] = Ain[i * 3 + j]; b[i] = bin[i]; }
    float maxn2 = 0.0f;
    for (int j = 0; j < 3; j++) { float s = 0.0f; for (int i = 0; i < 5; i++) s += A[i][j] * A[i][j]; if (s > maxn2) maxn2 = s; }
    float mx = sqrtf(maxn2) * FLT_EPSILON;
    float thr_helper = (mx * mx) / 5.0f;
    int nonzero = 3;
    for (int k = 0; k < 3; k++) {
        int piv = k; float best = -1.0f;
        for (int j = k; j < 3; j++) { float s = 0.0f; for (int i = k; i < 5; i++) s += A[i][j] * A[i][j]; if (s > best) { best = s; piv = j; } }
        if (nonzero == 3 && best < thr_helper * (float)(5 - k)) nonzero = k;
        if (piv != k) {
            for (int i = 0; i < 5; i++) { float t = A[i][k]; A[i][k] = A[i][piv]; A[i][piv] = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        float nrm = sqrtf(best);
        if (nrm == 0.0f) continue;
        float alpha = (A[k][k] >= 0.0f) ? -nrm : nrm;
        float v[5] = { 0, 0, 0, 0, 0 };
        for (int i = k; i < 5; i++) v[i] = A[i][k];
        v[k] = v[k] - alpha;
        float vn2 = 0.0f;
        for (int i = k; i < 5; i++) vn2 += v[i] * v[i];
        if (vn2 == 0.0f) continue;
        for (int j = k; j < 3; j++) {
            float dot = 0.0f;
            for (int i = k; i < 5; i++) dot += v[i] * A[i][j];
            float f = (2.0f * dot) / vn2;
            for (int i = k; i < 5; i++) A[i][j] = A[i][j] - f * v[i];
        }
        float dot = 0.0f;
        for (int i = k; i < 5; i++) dot += v[i] * b[i];
        float f = (2.0f * dot) / vn2;
        for (int i = k; i < 5; i++) b[i] = b[i] - f * v[i];
    }
    float y[3] = { 0, 0, 0 };
    for (int i = nonzero - 1; i >= 0; i--) {
        float s = b[i];
        for (int j = i + 1; j < nonzero; j++) s = s - A[i][j] * y[j];
        y[i] = s / A[i][i];
    }
    for (int i = 0; i < 3; i++) x[perm[i]] = y[i];
}

/* corner: line through the 5 neighbours' principal direction. returns 1 if kept */
int orc_map_edge_coeff(const orc_pt *sel, const orc_pt *nb /*5*/, float *coeff)
{
    float vx = 0.0f, vy = 0.0f, vz = 0.0f;
    for (int j = 0; j < 5; j++) { vx += nb[j].x; vy += nb[j].y; vz += nb[j].z; }
    vx = vx / 5.0f; vy = vy / 5.0f; vz = vz / 5.0f;
    float a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
    for (int j = 0; j < 5; j++) {
        float ax = nb[j].x - vx, ay = nb[j].y - vy, az = nb[j].z - vz;
        a00 += ax * ax; a10 += ax * ay; a20 += ax * az; a11 += ay * ay; a21 += ay * az; a22 += az * az;
    }
    float M[9];
    M[0] = a00 / 5.0f; M[4] = a11 / 5.0f; M[8] = a22 / 5.0f;
    M[3] = M[1] = a10 / 5.0f; M[6] = M[2] = a20 / 5.0f; M[7] = M[5] = a21 / 5.0f;
    float ev[3], evec[9];
    orc_eig3_jacobi(M, ev, evec);
    if (!(ev[2] > 3.0f * ev[1])) return 0;
    float x0 = sel->x, y0 = sel->y, z0 = sel->z;
    float x1 = (float)((double)vx + 0.1 * (double)evec[6]), y1 = (float)((double)vy + 0.1 * (double)evec[7]), z1 = (float)((double)vz + 0.1 * (double)evec[8]);
    float x2 = (float)((double)vx - 0.1 * (double)evec[6]), y2 = (float)((double)vy - 0.1 * (double)evec[7]), z2 = (float)((double)vz - 0.1 * (double)evec[8]);
    float m1 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    float m2 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    float m3 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    float a012 = sqrtf(m1 * m1 + m2 * m2 + m3 * m3);
    float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    float la = ((y1 - y2) * m1 + (z1 - z2) * m2) / a012 / l12;
    float lb = -((x1 - x2) * m1 - (z1 - z2) * m3) / a012 / l12;
    float lc = -((x1 - x2) * m2 + (y1 - y2) * m3) / a012 / l12;
    float ld2 = a012 / l12;
    float s = 1.0f - 0.9f * fabsf(ld2);
    coeff[0] = s * la; coeff[1] = s * lb; coeff[2] = s * lc; coeff[3] = s * ld2;
    return ((double)s > 0.1) ? 1 : 0;
}

int orc_map_plane_coeff(const orc_pt *sel, const orc_pt *nb /*5*/, float *coeff)
{
    float A0[15], B0[5] = { -1.0f, -1.0f, -1.0f, -1.0f, -1.0f }, X0[3];
    for (int j = 0; j < 5; j++) { A0[j * 3] = nb[j].x; A0[j * 3 + 1] = nb[j].y; A0[j * 3 + 2] = nb[j].z; }
    orc_lstsq53(A0, B0, X0);
    float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1.0f;
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa = pa / ps; pb = pb / ps; pc = pc / ps; pd = pd / ps;
    for (int j = 0; j < 5; j++)
        if ((double)fabsf(pa * nb[j].x + pb * nb[j].y + pc * nb[j].z + pd) > 0.2) return 0;
    float pd2 = pa * sel->x + pb * sel->y + pc * sel->z + pd;
    float s = 1.0f - 0.9f * fabsf(pd2) / sqrtf(sqrtf(sel->x * sel->x + sel->y * sel->y + sel->z * sel->z));
    coeff[0] = s * pa; coeff[1] = s * pb; coeff[2] = s * pc; coeff[3] = s * pd2;
    return ((double)s > 0.1) ? 1 : 0;
}

static void jacobian_row_map(const float *trig, const orc_pt *ori, const float *coeff, float *row, float *bval)
{
    float srx = trig[0], crx = trig[1], sry = trig[2], cry = trig[3], srz = trig[4], crz = trig[5];
    float x = ori->x, y = ori->y, z = ori->z;
    float cx_ = coeff[0], cy_ = coeff[1], cz_ = coeff[2];
    float arx = (crx * sry * srz * x + crx * crz * sry * y - srx * sry * z) * cx_
              + (-srx * srz * x - crz * srx * y - crx * z) * cy_
              + (crx * cry * srz * x + crx * cry * crz * y - cry * srx * z) * cz_;
    float ary = ((cry * srx * srz - crz * sry) * x + (sry * srz + cry * crz * srx) * y + crx * cry * z) * cx_
              + ((-cry * crz - srx * sry * srz) * x + (cry * srz - crz * srx * sry) * y - crx * sry * z) * cz_;
    float arz = ((crz * srx * sry - cry * srz) * x + (-cry * crz - srx * sry * srz) * y) * cx_
              + (crx * crz * x - crx * srz * y) * cy_
              + ((sry * srz + cry * crz * srx) * x + (crz * sry - cry * srx * srz) * y) * cz_;
    row[0] = arx; row[1] = ary; row[2] = arz; row[3] = cx_; row[4] = cy_; row[5] = cz_;
    *bval = -coeff[3];
}

void orc_map_jacobian_row(const float *T, const orc_pt *ori, const float *coeff, float *row, float *bval)
{
    float trig[6];
    orc_sincosf(T[0], &trig[0], &trig[1]); orc_sincosf(T[1], &trig[2], &trig[3]); orc_sincosf(T[2], &trig[4], &trig[5]);
    jacobian_row_map(trig, ori, coeff, row, bval);
}

void orc_point_to_map(const float *T, const orc_pt *in, int n, orc_pt *out)
{
    float trig[6];
    orc_sincosf(T[0], &trig[0], &trig[1]); orc_sincosf(T[1], &trig[2], &trig[3]); orc_sincosf(T[2], &trig[4], &trig[5]);
    for (int i = 0; i < n; i++) to_map(T, trig, &in[i], &out[i]);
}

static void fill_terms(const float *row, float bval, float w, float *t)
{
    int e = 0;
    for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) t[e++] = row[a] * row[b];
    for (int a = 0; a < 6; a++) t[e++] = row[a] * bval;
    t[e++] = w * w;
}

static void mapping_register_impl(const orc_config *c,
                          const orc_pt *corner_q, int n_cq, const orc_pt *surf_q, int n_sq,
                          const orc_pt *corner_map, int n_cm, const orc_pt *surf_map, int n_sm,
                          const float *seed, int use_kdtree, const orc_kdtree *kc_in, const orc_kdtree *ks_in,
                          orc_reg_result *res, int *trace_idx, float *trace_T);

void orc_mapping_register(const orc_config *c,
                          const orc_pt *corner_q, int n_cq, const orc_pt *surf_q, int n_sq,
                          const orc_pt *corner_map, int n_cm, const orc_pt *surf_map, int n_sm,
                          const float *seed, int use_kdtree,
                          orc_reg_result *res, int *trace_idx, float *trace_T)
{
    mapping_register_impl(c, corner_q, n_cq, surf_q, n_sq, corner_map, n_cm, surf_map, n_sm, seed, use_kdtree, NULL, NULL, res, trace_idx, trace_T);
}

/* same, with prebuilt kd-trees over the map clouds (CPU baseline: the trees are reused across frames) */
void orc_mapping_register_trees(const orc_config *c,
                                const orc_pt *corner_q, int n_cq, const orc_pt *surf_q, int n_sq,
                                const orc_pt *corner_map, int n_cm, const orc_kdtree *kc,
                                const orc_pt *surf_map, int n_sm, const orc_kdtree *ks,
                                const float *seed, orc_reg_result *res)
{
    mapping_register_impl(c, corner_q, n_cq, surf_q, n_sq, corner_map, n_cm, surf_map, n_sm, seed, 1, kc, ks, res, NULL, NULL);
}

static void mapping_register_impl(const orc_config *c,
                          const orc_pt *corner_q, int n_cq, const orc_pt *surf_q, int n_sq,
                          const orc_pt *corner_map, int n_cm, const orc_pt *surf_map, int n_sm,
                          const float *seed, int use_kdtree, const orc_kdtree *kc_in, const orc_kdtree *ks_in,
                          orc_reg_result *res, int *trace_idx, float *trace_T)
{
    memset(res, 0, sizeof(*res));
    float T[6];
    for (int a = 0; a < 6; a++) { T[a] = seed[a]; res->P[a * 7] = 1.0f; }
    res->status = 1;
    int Q = n_cq + n_sq;
    int it = 0;
    if (!(n_cm > 10 && n_sm > 100)) {     /* upstream early return */
        for (int a = 0; a < 6; a++) res->transform[a] = T[a];
        return;
    }
    orc_kdtree *kc_own = (use_kdtree && !kc_in) ? orc_kdtree_build(corner_map, n_cm) : NULL;
    orc_kdtree *ks_own = (use_kdtree && !ks_in) ? orc_kdtree_build(surf_map, n_sm) : NULL;
    const orc_kdtree *kc = kc_in ? kc_in : kc_own, *ks = ks_in ? ks_in : ks_own;
    float *terms = (float *)malloc(sizeof(float) * (size_t)(Q + 1) * NTERM);
    for (; it < c->map_max_iterations; it++) {
        float trig[6];
        orc_sincosf(T[0], &trig[0], &trig[1]); orc_sincosf(T[1], &trig[2], &trig[3]); orc_sincosf(T[2], &trig[4], &trig[5]);
        memset(terms, 0, sizeof(float) * (size_t)Q * NTERM);
        int n_edge = 0, n_plane = 0;
        for (int i = 0; i < Q; i++) {
            int is_corner = i < n_cq;
            const orc_pt *ori = is_corner ? &corner_q[i] : &surf_q[i - n_cq];
            const orc_pt *map = is_corner ? corner_map : surf_map;
            orc_pt sel; to_map(T, trig, ori, &sel);
            int idx[5]; float d2[5];
            if (use_kdtree) orc_kdtree_knn(is_corner ? kc : ks, &sel, 1, 5, idx, d2);
            else orc_knn_brute(map, is_corner ? n_cm : n_sm, &sel, 1, 5, idx, d2);
            if (trace_idx && it == 0) for (int j = 0; j < 5; j++) trace_idx[i * 5 + j] = (d2[4] < 1.0f) ? idx[j] : -1;
            if (!(d2[4] < 1.0f)) continue;
            orc_pt nb[5];
            for (int j = 0; j < 5; j++) nb[j] = map[idx[j]];
            float coeff[4], row[6], bval;
            int keep = is_corner ? orc_map_edge_coeff(&sel, nb, coeff) : orc_map_plane_coeff(&sel, nb, coeff);
            if (!keep) continue;
            jacobian_row_map(trig, ori, coeff, row, &bval);
            fill_terms(row, bval, coeff[3], terms + (size_t)i * NTERM);
            if (is_corner) n_edge++; else n_plane++;
        }
        if (trace_T) for (int a = 0; a < 6; a++) trace_T[it * 6 + a] = T[a];
        if (n_edge + n_plane < 50) continue;
        float total[NTERM];
        orc_reduce_r1(terms, Q, NTERM, total);
        res->n_corr_edge = n_edge; res->n_corr_plane = n_plane;
        int conv = 0;
        orc_gn_update(total, it, c->map_degen_eig, c->map_delta_t_abort, c->map_delta_r_abort, T, res, &conv, 1);
        orc_finish_result(c, total, n_edge + n_plane, res);
        res->status = 0;
        if (trace_T) for (int a = 0; a < 6; a++) trace_T[it * 6 + a] = T[a];
        if (conv) { it++; break; }
    }
    res->iterations = it;
    for (int a = 0; a < 6; a++) res->transform[a] = T[a];
    orc_kdtree_free(kc_own); orc_kdtree_free(ks_own);
    free(terms);
}
