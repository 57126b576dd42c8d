/* TEST INFRASTRUCTURE -- CPU oracle (see vlo_oracle.h; PARITY UNPINNED: the algorithm lives in the
 * un-vendored `loam` dependency, gtsam_fusion/package.xml:26, README.md:21-25).
 *
 * Restates BasicLaserOdometry::process of the LOAM fork, driven by
 * the reference's knobs odomMaxIterations/odomDeltaTAbort/odomDeltaRAbort/odomDegenEigVal
 * (gtsam_fusion/config/carla/loam_params.yaml:36-39).  Output feeds the reference's filter
 * (degerate_odometry_filter.cpp:30-46) through OptStatus.hessian.
 *
 * Frozen choices (upstream leaves them to Eigen / the compiler / the fork):
 *  R1  AtA / AtB summation order: terms indexed by query (sharp queries first, then flat; a
 *      rejected query contributes zeros); three-level blocked summation (32 queries, 32 blocks,
 *      then sequential), see orc_reduce_r1.  float32 throughout, no contraction.
 *  R2  P = sum over kept eigenvectors v v^T (== upstream matV^-1 matV2 for orthonormal V).
 *  R3  OptStatus.hessian = AtA of the LAST linearisation, native LOAM order (rx ry rz tx ty tz).
 *  R4  cov = sigma^2 (AtA)^-1, sigma^2 = sum (s d)^2 / (n - 6), float64.
 *  R5  sin/cos through orc_sincosf (detmath.h).
 */
#include "vlo_oracle.h"
#include "detmath.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_PI 3.14159265358979323846
#define NTERM 28   /* 21 upper-tri AtA + 6 AtB + 1 sum of squared weighted residuals */

static inline float sqd(const orc_pt *a, const orc_pt *b)
{
    float dx = a->x - b->x, dy = a->y - b->y, dz = a->z - b->z;
    return (dx * dx + dy * dy) + dz * dz;
}

static inline void rot_x(float *y, float *z, float c, float s) { float y0 = *y; *y = c * y0 - s * *z; *z = s * y0 + c * *z; }
static inline void rot_y(float *x, float *z, float c, float s) { float x0 = *x; *x = c * x0 + s * *z; *z = c * *z - s * x0; }
static inline void rot_z(float *x, float *y, float c, float s) { float x0 = *x; *x = c * x0 - s * *y; *y = s * x0 + c * *y; }

static inline float point_s(const orc_config *c, float w)
{
    if (!c->deskew) return 1.0f;
    return (1.0f / c->scan_period) * (w - (float)(int)w);
}

static void to_start(const orc_config *c, const float *T, const orc_pt *pi, orc_pt *po)
{
    float s = point_s(c, pi->w);
    float x = pi->x - s * T[3], y = pi->y - s * T[4], z = pi->z - s * T[5];
    float sx, cx, sy, cy, sz, cz;
    orc_sincosf(-s * T[0], &sx, &cx);
    orc_sincosf(-s * T[1], &sy, &cy);
    orc_sincosf(-s * T[2], &sz, &cz);
    rot_z(&x, &y, cz, sz);
    rot_x(&y, &z, cx, sx);
    rot_y(&x, &z, cy, sy);
    po->x = x; po->y = y; po->z = z; po->w = pi->w;
}

void orc_transform_to_start(const orc_config *c, const float *T, const orc_pt *in, int n, orc_pt *out)
{
    for (int i = 0; i < n; i++) to_start(c, T, &in[i], &out[i]);
}

void orc_transform_to_end(const orc_config *c, const float *T, orc_pt *cloud, int n)
{
    float sx, cx, sy, cy, sz, cz;
    orc_sincosf(T[0], &sx, &cx);
    orc_sincosf(T[1], &sy, &cy);
    orc_sincosf(T[2], &sz, &cz);
    for (int i = 0; i < n; i++) {
        orc_pt p;
        to_start(c, T, &cloud[i], &p);
        float x = p.x, y = p.y, z = p.z;
        /* rotateYXZ(point, ry, rx, rz) */
        rot_y(&x, &z, cy, sy);
        rot_x(&y, &z, cx, sx);
        rot_z(&x, &y, cz, sz);
        cloud[i].x = x + T[3]; cloud[i].y = y + T[4]; cloud[i].z = z + T[5];
        cloud[i].w = (float)(int)cloud[i].w;
    }
}

/* partner search shared by corner / surface association; semantics of the upstream loops:
 * forward j = ind+1.. (ascending, strict <), then backward j = ind-1.. (descending, strict <),
 * sharing the running minimum. */
static void partners_corner(const orc_pt *cl, int n, int fwd_bound, const orc_pt *q, int ind, int *ind2)
{
    int scan = (int)cl[ind].w;
    float best = 25.0f; int b = -1;
    for (int j = ind + 1; j < fwd_bound; j++) {
        if ((double)(int)cl[j].w > (double)scan + 2.5) break;
        float d = sqd(&cl[j], q);
        if ((int)cl[j].w > scan) { if (d < best) { best = d; b = j; } }
    }
    for (int j = ind - 1; j >= 0; j--) {
        if ((double)(int)cl[j].w < (double)scan - 2.5) break;
        float d = sqd(&cl[j], q);
        if ((int)cl[j].w < scan) { if (d < best) { best = d; b = j; } }
    }
    (void)n;
    *ind2 = b;
}

static void partners_surf(const orc_pt *cl, int n, int fwd_bound, const orc_pt *q, int ind, int *ind2, int *ind3)
{
    int scan = (int)cl[ind].w;
    float best2 = 25.0f, best3 = 25.0f; int b2 = -1, b3 = -1;
    for (int j = ind + 1; j < fwd_bound; j++) {
        if ((double)(int)cl[j].w > (double)scan + 2.5) break;
        float d = sqd(&cl[j], q);
        if ((int)cl[j].w <= scan) { if (d < best2) { best2 = d; b2 = j; } }
        else                      { if (d < best3) { best3 = d; b3 = j; } }
    }
    for (int j = ind - 1; j >= 0; j--) {
        if ((double)(int)cl[j].w < (double)scan - 2.5) break;
        float d = sqd(&cl[j], q);
        if ((int)cl[j].w >= scan) { if (d < best2) { best2 = d; b2 = j; } }
        else                      { if (d < best3) { best3 = d; b3 = j; } }
    }
    (void)n;
    *ind2 = b2; *ind3 = b3;
}

static void associate(const orc_config *c, const float *T,
                      const orc_pt *cur_sharp, int n_sharp, const orc_pt *cur_flat, int n_flat,
                      const orc_pt *last_corner, int n_lc, const orc_kdtree *kc,
                      const orc_pt *last_surf, int n_ls, const orc_kdtree *ks,
                      int *corner_idx, int *surf_idx)
{
    int fb_c = c->odom_forward_bound_quirk ? (n_sharp < n_lc ? n_sharp : n_lc) : n_lc;
    int fb_s = c->odom_forward_bound_quirk ? (n_flat < n_ls ? n_flat : n_ls) : n_ls;
    for (int i = 0; i < n_sharp; i++) {
        orc_pt sel; to_start(c, T, &cur_sharp[i], &sel);
        int ni = -1; float nd = INFINITY;
        if (kc) orc_kdtree_knn(kc, &sel, 1, 1, &ni, &nd); else orc_knn_brute(last_corner, n_lc, &sel, 1, 1, &ni, &nd);
        int i1 = -1, i2 = -1;
        if (ni >= 0 && nd < 25.0f) { i1 = ni; partners_corner(last_corner, n_lc, fb_c, &sel, i1, &i2); }
        corner_idx[2 * i] = i1; corner_idx[2 * i + 1] = i2;
    }
    for (int i = 0; i < n_flat; i++) {
        orc_pt sel; to_start(c, T, &cur_flat[i], &sel);
        int ni = -1; float nd = INFINITY;
        if (ks) orc_kdtree_knn(ks, &sel, 1, 1, &ni, &nd); else orc_knn_brute(last_surf, n_ls, &sel, 1, 1, &ni, &nd);
        int i1 = -1, i2 = -1, i3 = -1;
        if (ni >= 0 && nd < 25.0f) { i1 = ni; partners_surf(last_surf, n_ls, fb_s, &sel, i1, &i2, &i3); }
        surf_idx[3 * i] = i1; surf_idx[3 * i + 1] = i2; surf_idx[3 * i + 2] = i3;
    }
}

void orc_odometry_associate(const orc_config *c, const float *T,
                            const orc_pt *cur_sharp, int n_sharp, const orc_pt *cur_flat, int n_flat,
                            const orc_pt *last_corner, int n_lc, const orc_pt *last_surf, int n_ls, int use_kdtree,
                            int *corner_idx, int *surf_idx)
{
    orc_kdtree *kc = use_kdtree ? orc_kdtree_build(last_corner, n_lc) : NULL;
    orc_kdtree *ks = use_kdtree ? orc_kdtree_build(last_surf, n_ls) : NULL;
    associate(c, T, cur_sharp, n_sharp, cur_flat, n_flat, last_corner, n_lc, kc, last_surf, n_ls, ks, corner_idx, surf_idx);
    orc_kdtree_free(kc); orc_kdtree_free(ks);
}

/* R1 reduction of terms[Q][nterm]: three-level blocked summation, float32.
 *   level 1: blocks of 32 consecutive queries, each summed sequentially from its first element;
 *   level 2: blocks of 32 consecutive level-1 sums (1024 queries), summed sequentially;
 *   level 3: level-2 sums added sequentially. */
void orc_reduce_r1(const float *terms, int Q, int nterm, float *total)
{
    for (int e = 0; e < nterm; e++) {
        float l3 = 0.0f;
        for (int c2 = 0; c2 < Q; c2 += 1024) {
            float l2 = 0.0f;
            for (int c1 = c2; c1 < Q && c1 < c2 + 1024; c1 += 32) {
                float l1 = 0.0f;
                for (int i = c1; i < Q && i < c1 + 32; i++) l1 = l1 + terms[(size_t)i * nterm + e];
                l2 = l2 + l1;
            }
            l3 = l3 + l2;
        }
        total[e] = l3;
    }
}

static void jacobian_row_odom(const float *T, const float *trig, const orc_pt *ori, const float *coeff, float *row, float *bval)
{
    /* trig = srx crx sry cry srz crz  (s = 1 in upstream's Jacobian) */
    float srx = trig[0], crx = trig[1], sry = trig[2], cry = trig[3], srz = trig[4], crz = trig[5];
    float tx = T[3], ty = T[4], tz = T[5];
    float x = ori->x, y = ori->y, z = ori->z;
    float cx_ = coeff[0], cy_ = coeff[1], cz_ = coeff[2];
    float arx = (-crx * sry * srz * x + crx * crz * sry * y + srx * sry * z
                 + tx * crx * sry * srz - ty * crx * crz * sry - tz * srx * sry) * cx_
              + (srx * srz * x - crz * srx * y + crx * z
                 + ty * crz * srx - tz * crx - tx * srx * srz) * cy_
              + (crx * cry * srz * x - crx * cry * crz * y - cry * srx * z
                 + tz * cry * srx + ty * crx * cry * crz - tx * crx * cry * srz) * cz_;
    float ary = ((-crz * sry - cry * srx * srz) * x
                 + (cry * crz * srx - sry * srz) * y - crx * cry * z
                 + tx * (crz * sry + cry * srx * srz) + ty * (sry * srz - cry * crz * srx)
                 + tz * crx * cry) * cx_
              + ((cry * crz - srx * sry * srz) * x
                 + (cry * srz + crz * srx * sry) * y - crx * sry * z
                 + tz * crx * sry - ty * (cry * srz + crz * srx * sry)
                 - tx * (cry * crz - srx * sry * srz)) * cz_;
    float arz = ((-cry * srz - crz * srx * sry) * x + (cry * crz - srx * sry * srz) * y
                 + tx * (cry * srz + crz * srx * sry) - ty * (cry * crz - srx * sry * srz)) * cx_
              + (-crx * crz * x - crx * srz * y
                 + ty * crx * srz + tx * crx * crz) * cy_
              + ((cry * crz * srx - sry * srz) * x + (crz * sry + cry * srx * srz) * y
                 + tx * (sry * srz - cry * crz * srx) - ty * (crz * sry + cry * srx * srz)) * cz_;
    float atx = -(cry * crz - srx * sry * srz) * cx_ + crx * srz * cy_ - (crz * sry + cry * srx * srz) * cz_;
    float aty = -(cry * srz + crz * srx * sry) * cx_ - crx * crz * cy_ - (sry * srz - cry * crz * srx) * cz_;
    float atz = crx * sry * cx_ - srx * cy_ - crx * cry * cz_;
    row[0] = arx; row[1] = ary; row[2] = arz; row[3] = atx; row[4] = aty; row[5] = atz;
    *bval = (float)(-0.05 * (double)coeff[3]);
}

/* exposed for tests (finite-difference check of the Jacobian) */
void orc_odom_jacobian_row(const float *T, const orc_pt *ori, const float *coeff, float *row, float *bval)
{
    float trig[6];
    orc_sincosf(T[0], &trig[0], &trig[1]);
    orc_sincosf(T[1], &trig[2], &trig[3]);
    orc_sincosf(T[2], &trig[4], &trig[5]);
    jacobian_row_odom(T, trig, ori, coeff, row, bval);
}

/* edge / plane coefficient (A.5). returns 1 if the correspondence is kept */
int orc_edge_coeff(const orc_pt *sel, const orc_pt *a, const orc_pt *b, int iter, float *coeff)
{
    float x0 = sel->x, y0 = sel->y, z0 = sel->z;
    float x1 = a->x, y1 = a->y, z1 = a->z;
    float x2 = b->x, y2 = b->y, z2 = b->z;
    float m1 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    float m2 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    float m3 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    float a012 = sqrtf(m1 * m1 + m2 * m2 + m3 * m3);
    float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    float la = ((y1 - y2) * m1 + (z1 - z2) * m2) / a012 / l12;
    float lb = -((x1 - x2) * m1 - (z1 - z2) * m3) / a012 / l12;
    float lc = -((x1 - x2) * m2 + (y1 - y2) * m3) / a012 / l12;
    float ld2 = a012 / l12;
    float s = 1.0f;
    if (iter >= 5) s = 1.0f - 1.8f * fabsf(ld2);
    coeff[0] = s * la; coeff[1] = s * lb; coeff[2] = s * lc; coeff[3] = s * ld2;
    return ((double)s > 0.1 && ld2 != 0.0f) ? 1 : 0;
}

int orc_plane_coeff(const orc_pt *sel, const orc_pt *t1, const orc_pt *t2, const orc_pt *t3, int iter, float *coeff)
{
    float pa = (t2->y - t1->y) * (t3->z - t1->z) - (t3->y - t1->y) * (t2->z - t1->z);
    float pb = (t2->z - t1->z) * (t3->x - t1->x) - (t3->z - t1->z) * (t2->x - t1->x);
    float pc = (t2->x - t1->x) * (t3->y - t1->y) - (t3->x - t1->x) * (t2->y - t1->y);
    float pd = -(pa * t1->x + pb * t1->y + pc * t1->z);
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa = pa / ps; pb = pb / ps; pc = pc / ps; pd = pd / ps;
    float pd2 = pa * sel->x + pb * sel->y + pc * sel->z + pd;
    float s = 1.0f;
    if (iter >= 5) {
        float dist = sqrtf(sel->x * sel->x + sel->y * sel->y + sel->z * sel->z);
        s = 1.0f - 1.8f * fabsf(pd2) / sqrtf(dist);
    }
    coeff[0] = s * pa; coeff[1] = s * pb; coeff[2] = s * pc; coeff[3] = s * pd2;
    return ((double)s > 0.1 && pd2 != 0.0f) ? 1 : 0;
}

static void fill_terms(const float *row, float bval, float w, float *t)
{
    int e = 0;
    for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) t[e++] = row[a] * row[b];
    for (int a = 0; a < 6; a++) t[e++] = row[a] * bval;
    t[e++] = w * w;
}

/* float64 inverse via Gauss-Jordan with partial pivoting; returns 0 if singular */
static int inv6d(const double *A, double *Ai)
{
    double M[6][12];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { M[i][j] = A[i * 6 + j]; M[i][6 + j] = (i == j); }
    for (int k = 0; k < 6; k++) {
        int p = k; double mx = fabs(M[k][k]);
        for (int i = k + 1; i < 6; i++) if (fabs(M[i][k]) > mx) { mx = fabs(M[i][k]); p = i; }
        if (mx == 0.0) return 0;
        if (p != k) for (int j = 0; j < 12; j++) { double t = M[k][j]; M[k][j] = M[p][j]; M[p][j] = t; }
        double d = M[k][k];
        for (int j = 0; j < 12; j++) M[k][j] /= d;
        for (int i = 0; i < 6; i++) if (i != k) { double f = M[i][k]; for (int j = 0; j < 12; j++) M[i][j] -= f * M[k][j]; }
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Ai[i * 6 + j] = M[i][6 + j];
    return 1;
}

void orc_finish_result(const orc_config *c, const float *total, int n_corr, orc_reg_result *res)
{
    /* unpack H, D-opt gate (degerate_odometry_filter.cpp:30-46), covariance (R4) */
    int e = 0;
    for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) { res->hessian[a * 6 + b] = total[e]; res->hessian[b * 6 + a] = total[e]; e++; }
    res->pass_dopt = orc_dopt_gate(res->hessian, (double)c->dopt_rot_threshold, (double)c->dopt_trans_threshold,
                                   &res->logdet_rot, &res->logdet_trans);
    double Hd[36], Hi[36];
    for (int i = 0; i < 36; i++) Hd[i] = (double)res->hessian[i];
    double dof = n_corr > 6 ? (double)(n_corr - 6) : 1.0;
    double sigma2 = (double)total[27] / dof;
    if (inv6d(Hd, Hi)) for (int i = 0; i < 36; i++) res->cov[i] = sigma2 * Hi[i];
    else for (int i = 0; i < 36; i++) res->cov[i] = NAN;
}

void orc_gn_update(const float *total, int iter, float degen_thr, float dT_abort, float dR_abort,
                   float *T, orc_reg_result *res, int *converged, int map_order)
{
    float H[36], g[6], x[6];
    int e = 0;
    for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) { H[a * 6 + b] = total[e]; H[b * 6 + a] = total[e]; e++; }
    for (int a = 0; a < 6; a++) g[a] = total[21 + a];
    orc_solve6_colpiv_qr(H, g, x);
    if (iter == 0) {
        res->is_degenerate = orc_degeneracy(H, degen_thr, res->eig, res->P);
    }
    if (res->is_degenerate) {
        float x2[6];
        for (int a = 0; a < 6; a++) x2[a] = x[a];
        for (int a = 0; a < 6; a++) {
            float s = 0.0f;
            for (int b = 0; b < 6; b++) s += res->P[a * 6 + b] * x2[b];
            x[a] = s;
        }
    }
    for (int a = 0; a < 6; a++) {
        T[a] = T[a] + x[a];
        if (!isfinite(T[a])) T[a] = 0.0f;
    }
    (void)map_order;
    double r0 = (double)(float)((double)x[0] * 180.0 / ORC_PI);
    double r1 = (double)(float)((double)x[1] * 180.0 / ORC_PI);
    double r2 = (double)(float)((double)x[2] * 180.0 / ORC_PI);
    float deltaR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    double t0 = (double)(x[3] * 100.0f), t1 = (double)(x[4] * 100.0f), t2 = (double)(x[5] * 100.0f);
    float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    *converged = (deltaR < dR_abort && deltaT < dT_abort) ? 1 : 0;
}

void orc_odometry_register(const orc_config *c,
                           const orc_pt *cur_sharp, int n_sharp, const orc_pt *cur_flat, int n_flat,
                           const orc_pt *last_corner, int n_lc, const int *lc_ring_start,
                           const orc_pt *last_surf, int n_ls, const int *ls_ring_start,
                           const float *seed, int use_kdtree,
                           orc_reg_result *res, int *trace_idx, float *trace_T)
{
    (void)lc_ring_start; (void)ls_ring_start;
    memset(res, 0, sizeof(*res));
    float T[6];
    for (int a = 0; a < 6; a++) T[a] = seed[a];
    for (int a = 0; a < 6; a++) res->P[a * 7] = 1.0f;
    int Q = n_sharp + n_flat;
    int *cidx = (int *)malloc(sizeof(int) * (size_t)(2 * n_sharp + 1));
    int *sidx = (int *)malloc(sizeof(int) * (size_t)(3 * n_flat + 1));
    float *terms = (float *)malloc(sizeof(float) * (size_t)(Q + 1) * NTERM);
    for (int i = 0; i < 2 * n_sharp; i++) cidx[i] = -1;
    for (int i = 0; i < 3 * n_flat; i++) sidx[i] = -1;
    orc_kdtree *kc = use_kdtree ? orc_kdtree_build(last_corner, n_lc) : NULL;
    orc_kdtree *ks = use_kdtree ? orc_kdtree_build(last_surf, n_ls) : NULL;
    int trace_off = 0;
    res->status = 1;
    /* upstream only optimises when the last clouds are big enough */
    int enough = (n_lc > 10 && n_ls > 100);
    int it = 0;
    for (; enough && it < c->odom_max_iterations; it++) {
        if (it % 5 == 0) {
            associate(c, T, cur_sharp, n_sharp, cur_flat, n_flat, last_corner, n_lc, kc, last_surf, n_ls, ks, cidx, sidx);
            if (trace_idx) {
                memcpy(trace_idx + trace_off, cidx, sizeof(int) * (size_t)(2 * n_sharp)); trace_off += 2 * n_sharp;
                memcpy(trace_idx + trace_off, sidx, sizeof(int) * (size_t)(3 * n_flat)); trace_off += 3 * n_flat;
            }
        }
        float trig[6];
        orc_sincosf(T[0], &trig[0], &trig[1]);
        orc_sincosf(T[1], &trig[2], &trig[3]);
        orc_sincosf(T[2], &trig[4], &trig[5]);
        int n_edge = 0, n_plane = 0;
        memset(terms, 0, sizeof(float) * (size_t)Q * NTERM);
        for (int i = 0; i < n_sharp; i++) {
            if (cidx[2 * i + 1] < 0) continue;
            orc_pt sel; to_start(c, T, &cur_sharp[i], &sel);
            float coeff[4], row[6], bval;
            if (!orc_edge_coeff(&sel, &last_corner[cidx[2 * i]], &last_corner[cidx[2 * i + 1]], it, coeff)) continue;
            jacobian_row_odom(T, trig, &cur_sharp[i], coeff, row, &bval);
            fill_terms(row, bval, coeff[3], terms + (size_t)i * NTERM);
            n_edge++;
        }
        for (int i = 0; i < n_flat; i++) {
            if (sidx[3 * i + 1] < 0 || sidx[3 * i + 2] < 0) continue;
            orc_pt sel; to_start(c, T, &cur_flat[i], &sel);
            float coeff[4], row[6], bval;
            if (!orc_plane_coeff(&sel, &last_surf[sidx[3 * i]], &last_surf[sidx[3 * i + 1]], &last_surf[sidx[3 * i + 2]], it, coeff)) continue;
            jacobian_row_odom(T, trig, &cur_flat[i], coeff, row, &bval);
            fill_terms(row, bval, coeff[3], terms + (size_t)(n_sharp + i) * NTERM);
            n_plane++;
        }
        if (trace_T) { /* filled after update below; pre-fill with current T for skipped iterations */
            for (int a = 0; a < 6; a++) trace_T[it * 6 + a] = T[a];
        }
        if (n_edge + n_plane < 10) continue;
        float total[NTERM];
        orc_reduce_r1(terms, Q, NTERM, total);
        res->n_corr_edge = n_edge; res->n_corr_plane = n_plane;
        int conv = 0;
        orc_gn_update(total, it, c->odom_degen_eig, c->odom_delta_t_abort, c->odom_delta_r_abort, T, res, &conv, 0);
        orc_finish_result(c, total, n_edge + n_plane, res);
        res->status = 0;
        if (trace_T) for (int a = 0; a < 6; a++) trace_T[it * 6 + a] = T[a];
        if (conv) { it++; break; }
    }
    res->iterations = it;
    for (int a = 0; a < 6; a++) res->transform[a] = T[a];
    orc_kdtree_free(kc); orc_kdtree_free(ks);
    free(cidx); free(sidx); free(terms);
}

/* pose accumulation: transformSum <- transformSum (+) T.  R = Ry Rx Rz (rotateZXY).
 * Exact composition R_sum' = R_sum R_T^-1, p_sum' = p_sum - R_sum' t (upstream's
 * accumulateRotation is the small-angle version of the same); float64. */
static void euler_to_R(double rx, double ry, double rz, double R[3][3])
{
    double sx = sin(rx), cx = cos(rx), sy = sin(ry), cy = cos(ry), sz = sin(rz), cz = cos(rz);
    R[0][0] = cy * cz + sy * sx * sz; R[0][1] = -cy * sz + sy * sx * cz; R[0][2] = sy * cx;
    R[1][0] = cx * sz;                R[1][1] = cx * cz;                 R[1][2] = -sx;
    R[2][0] = -sy * cz + cy * sx * sz; R[2][1] = sy * sz + cy * sx * cz; R[2][2] = cy * cx;
}

void orc_accumulate_pose(const float *sum_in, const float *T, float fudge, float *sum_out)
{
    double Rs[3][3], Rt[3][3], Rn[3][3];
    euler_to_R(sum_in[0], sum_in[1], sum_in[2], Rs);
    euler_to_R(T[0], (double)T[1] * fudge, T[2], Rt);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += Rs[i][k] * Rt[j][k];   /* Rs * Rt^T */
        Rn[i][j] = s;
    }
    double t[3] = { T[3], T[4], (double)T[5] * fudge };
    double rx = -asin(Rn[1][2]);
    double ry = atan2(Rn[0][2], Rn[2][2]);
    double rz = atan2(Rn[1][0], Rn[1][1]);
    sum_out[0] = (float)rx; sum_out[1] = (float)ry; sum_out[2] = (float)rz;
    for (int i = 0; i < 3; i++) {
        double s = 0; for (int k = 0; k < 3; k++) s += Rn[i][k] * t[k];
        sum_out[3 + i] = (float)((double)sum_in[3 + i] - s);
    }
}
