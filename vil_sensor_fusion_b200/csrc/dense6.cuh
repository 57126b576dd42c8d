// K4  6x6 float32 device routines of the Gauss-Newton step: column-pivoted Householder QR solve,
// single-warp round-robin Jacobi eigen-decomposition, degeneracy projection, D-optimality gate and
// covariance.  Replaces Eigen's colPivHouseholderQr / SelfAdjointEigenSolver in the `loam` nodelets
//  and folds in gtsam_fusion/src/degerate_odometry_filter.cpp:30-46.
// Operation orders are the ones frozen by the oracle (oracle/dense6.c) so results are bit-identical.
#pragma once
#include "vlo_internal.cuh"
#include <float.h>

// Fully unrolled with static indices so the 6x6 working set lives in registers (dynamic indexing
// would push it to local memory: round-1 profile of the online tick); the pivot swap and the
// rank-limited back substitution are predicated instead of indexed.  Same operations, same order.
__device__ inline void vlo_solve6_colpiv_qr(const float *Ain, const float *bin, float *x)
{
    float A[6][6], b[6];
    int perm[6];
    #pragma unroll
    for (int i = 0; i < 6; i++) {
        #pragma unroll
        for (int j = 0; j < 6; j++) A[i][j] = Ain[i * 6 + j];
        b[i] = bin[i]; perm[i] = i;
    }
    float maxn2 = 0.0f;
    #pragma unroll
    for (int j = 0; j < 6; j++) {
        float s = 0.0f;
        #pragma unroll
        for (int i = 0; i < 6; i++) s += A[i][j] * A[i][j];
        if (s > maxn2) maxn2 = s;
    }
    float mx = sqrtf(maxn2) * FLT_EPSILON;
    float thr_helper = (mx * mx) / 6.0f;
    int nonzero = 6;
    #pragma unroll
    for (int k = 0; k < 6; k++) {
        int piv = k; float best = -1.0f;
        #pragma unroll
        for (int j = k; j < 6; j++) {
            float s = 0.0f;
            #pragma unroll
            for (int i = k; i < 6; i++) s += A[i][j] * A[i][j];
            if (s > best) { best = s; piv = j; }
        }
        if (nonzero == 6 && best < thr_helper * (float)(6 - k)) nonzero = k;
        #pragma unroll
        for (int j = k + 1; j < 6; j++) {
            if (piv == j) {
                #pragma unroll
                for (int i = 0; i < 6; i++) { float t = A[i][k]; A[i][k] = A[i][j]; A[i][j] = t; }
                int t = perm[k]; perm[k] = perm[j]; perm[j] = t;
            }
        }
        float nrm = sqrtf(best);
        if (nrm != 0.0f) {
            float alpha = (A[k][k] >= 0.0f) ? -nrm : nrm;
            float v[6];
            #pragma unroll
            for (int i = 0; i < 6; i++) v[i] = (i >= k) ? A[i][k] : 0.0f;
            v[k] = v[k] - alpha;
            float vn2 = 0.0f;
            #pragma unroll
            for (int i = k; i < 6; i++) vn2 += v[i] * v[i];
            if (vn2 != 0.0f) {
                #pragma unroll
                for (int j = k; j < 6; j++) {
                    float dot = 0.0f;
                    #pragma unroll
                    for (int i = k; i < 6; i++) dot += v[i] * A[i][j];
                    float f = (2.0f * dot) / vn2;
                    #pragma unroll
                    for (int i = k; i < 6; i++) A[i][j] = A[i][j] - f * v[i];
                }
                float dot = 0.0f;
                #pragma unroll
                for (int i = k; i < 6; i++) dot += v[i] * b[i];
                float f = (2.0f * dot) / vn2;
                #pragma unroll
                for (int i = k; i < 6; i++) b[i] = b[i] - f * v[i];
            }
        }
    }
    float y[6];
    #pragma unroll
    for (int i = 0; i < 6; i++) y[i] = 0.0f;
    #pragma unroll
    for (int i = 5; i >= 0; i--) {
        if (i < nonzero) {
            float s = b[i];
            #pragma unroll
            for (int j = i + 1; j < 6; j++) if (j < nonzero) s = s - A[i][j] * y[j];
            y[i] = s / A[i][i];
        }
    }
    #pragma unroll
    for (int i = 0; i < 6; i++) x[perm[i]] = y[i];
}

// tournament order: 5 rounds x 3 disjoint pairs
__constant__ int VLO_JROUND[5][3][2] = {
    { {0, 5}, {1, 4}, {2, 3} }, { {0, 4}, {3, 5}, {1, 2} }, { {0, 3}, {2, 4}, {1, 5} },
    { {0, 2}, {1, 3}, {4, 5} }, { {0, 1}, {2, 5}, {3, 4} },
};
#define VLO_JACOBI_SWEEPS 8

// Single-warp Jacobi: A, V live in shared memory (36 floats each); lanes 0..2 compute the three
// rotations of a round, lanes 0..17 apply them ((pair, row) tasks), matrix re-symmetrised per round.
// On return (after __syncwarp) eval[6] ascending, evec row i = eigenvector i.  All 32 lanes must call.
__device__ inline void vlo_eig6_jacobi_warp(float *A, float *V, float *cs_sn /*6 floats smem*/, float *eval, float *evec, int lane)
{
    for (int k = lane; k < 36; k += 32) V[k] = (k / 6 == k % 6) ? 1.0f : 0.0f;
    __syncwarp();
    if (lane < 15) {   // symmetrise from the upper triangle
        int i = 0, r = lane; while (r >= 5 - i) { r -= 5 - i; i++; }
        int j = i + 1 + r;
        A[j * 6 + i] = A[i * 6 + j];
    }
    __syncwarp();
    for (int sweep = 0; sweep < VLO_JACOBI_SWEEPS; sweep++) {
        for (int r = 0; r < 5; r++) {
            if (lane < 3) {
                int p = VLO_JROUND[r][lane][0], q = VLO_JROUND[r][lane][1];
                float apq = A[p * 6 + q];
                float c = 1.0f, s = 0.0f;
                if (apq != 0.0f) {
                    float theta = (A[q * 6 + q] - A[p * 6 + p]) / (2.0f * apq);
                    float t = 1.0f / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
                    if (theta < 0.0f) t = -t;
                    c = 1.0f / sqrtf(t * t + 1.0f);
                    s = t * c;
                }
                cs_sn[lane * 2] = c; cs_sn[lane * 2 + 1] = s;
            }
            __syncwarp();
            int m = lane / 6, i = lane % 6;          // task (pair m, row/col i) for lanes 0..17
            int p = 0, q = 0; float c = 1.f, s = 0.f;
            if (lane < 18) { p = VLO_JROUND[r][m][0]; q = VLO_JROUND[r][m][1]; c = cs_sn[m * 2]; s = cs_sn[m * 2 + 1]; }
            if (lane < 18) {                          // columns: A <- A J
                float aip = A[i * 6 + p], aiq = A[i * 6 + q];
                A[i * 6 + p] = c * aip - s * aiq;
                A[i * 6 + q] = s * aip + c * aiq;
            }
            __syncwarp();
            if (lane < 18) {                          // rows: A <- J^T A
                float apj = A[p * 6 + i], aqj = A[q * 6 + i];
                A[p * 6 + i] = c * apj - s * aqj;
                A[q * 6 + i] = s * apj + c * aqj;
                float vip = V[i * 6 + p], viq = V[i * 6 + q];   // V <- V J
                V[i * 6 + p] = c * vip - s * viq;
                V[i * 6 + q] = s * vip + c * viq;
            }
            __syncwarp();
            if (lane < 15) {
                int ii = 0, rr = lane; while (rr >= 5 - ii) { rr -= 5 - ii; ii++; }
                int jj = ii + 1 + rr;
                A[jj * 6 + ii] = A[ii * 6 + jj];
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        int order[6] = { 0, 1, 2, 3, 4, 5 };
        for (int i = 1; i < 6; i++) {
            int v = order[i]; int j = i;
            while (j >= 1 && A[v * 6 + v] < A[order[j - 1] * 6 + order[j - 1]]) { order[j] = order[j - 1]; j--; }
            order[j] = v;
        }
        for (int i = 0; i < 6; i++) {
            eval[i] = A[order[i] * 6 + order[i]];
            for (int k = 0; k < 6; k++) evec[i * 6 + k] = V[k * 6 + order[i]];
        }
    }
    __syncwarp();
}

__device__ inline float vlo_det3(const float *H, int o)
{
#define M_(r, c) H[(o + (r)) * 6 + (o + (c))]
    float a = M_(0, 0) * (M_(1, 1) * M_(2, 2) - M_(1, 2) * M_(2, 1));
    float b = M_(0, 1) * (M_(1, 0) * M_(2, 2) - M_(1, 2) * M_(2, 0));
    float c = M_(0, 2) * (M_(1, 0) * M_(2, 1) - M_(1, 1) * M_(2, 0));
#undef M_
    return (a - b) + c;
}

// Shared-memory scratch of one Gauss-Newton problem (one CTA or one warp owns it)
struct GnScratch {
    float total[VLO_NTERM];
    float H[36];
    float A[36], V[36];
    float cs_sn[8];
    float eval[6], evec[36];
    float P[36];
    float T[6];
    float x[6];
    int   is_degenerate, converged, n_edge, n_plane, status, iterations;
};

// One GN update from the reduced totals (oracle: orc_gn_update).  Called by a full warp (warp 0 of
// the CTA); lane 0 does the scalar parts.  iter == 0 computes the degeneracy projection.
__device__ inline void vlo_gn_update_warp(GnScratch &S, int iter, float degen_thr, float dT_abort, float dR_abort, int lane)
{
    if (lane == 0) {
        int e = 0;
        for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) { S.H[a * 6 + b] = S.total[e]; S.H[b * 6 + a] = S.total[e]; e++; }
        float g[6];
        for (int a = 0; a < 6; a++) g[a] = S.total[21 + a];
        vlo_solve6_colpiv_qr(S.H, g, S.x);
    }
    __syncwarp();
    if (iter == 0) {
        for (int k = lane; k < 36; k += 32) S.A[k] = S.H[k];
        __syncwarp();
        vlo_eig6_jacobi_warp(S.A, S.V, S.cs_sn, S.eval, S.evec, lane);
        if (lane == 0) {
            int n_drop = 0;
            for (int i = 0; i < 6; i++) { if (S.eval[i] < degen_thr) n_drop++; else break; }
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) {
                    float s = 0.0f;
                    for (int i = n_drop; i < 6; i++) s += S.evec[i * 6 + a] * S.evec[i * 6 + b];
                    S.P[a * 6 + b] = s;
                }
            S.is_degenerate = n_drop > 0;
        }
        __syncwarp();
    }
    if (lane == 0) {
        float x[6];
        for (int a = 0; a < 6; a++) x[a] = S.x[a];
        if (S.is_degenerate) {
            for (int a = 0; a < 6; a++) {
                float s = 0.0f;
                for (int b = 0; b < 6; b++) s += S.P[a * 6 + b] * S.x[b];
                x[a] = s;
            }
        }
        for (int a = 0; a < 6; a++) {
            float t = S.T[a] + x[a];
            if (!isfinite(t)) t = 0.0f;
            S.T[a] = t;
        }
        double r0 = (double)(float)((double)x[0] * 180.0 / VLO_PI_D);
        double r1 = (double)(float)((double)x[1] * 180.0 / VLO_PI_D);
        double r2 = (double)(float)((double)x[2] * 180.0 / VLO_PI_D);
        float deltaR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
        double t0 = (double)(x[3] * 100.0f), t1 = (double)(x[4] * 100.0f), t2 = (double)(x[5] * 100.0f);
        float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
        S.converged = (deltaR < dR_abort && deltaT < dT_abort) ? 1 : 0;
    }
    __syncwarp();
}

// result record from the last linearisation's totals (oracle: orc_finish_result); one thread.
// The float64 covariance sigma^2 (AtA)^-1 is finished on the host right after the result record is
// copied back (vlo_finish_cov_host): the device only stages sum((s d)^2) in cov[0] and n in cov[1].
__device__ inline void vlo_finish_result(const GnScratch &S, float rot_thr, float trans_thr, vlo_result *res)
{
    int e = 0;
    #pragma unroll
    for (int a = 0; a < 6; a++)
        #pragma unroll
        for (int b = a; b < 6; b++) { res->hessian[a * 6 + b] = S.total[e]; res->hessian[b * 6 + a] = S.total[e]; e++; }
    float rot = logf(vlo_det3(res->hessian, 3));
    float trans = logf(vlo_det3(res->hessian, 0));
    res->logdet_rot = rot; res->logdet_trans = trans;
    res->pass_dopt = ((double)rot < (double)rot_thr || (double)trans < (double)trans_thr) ? 0 : 1;
    res->cov[0] = (double)S.total[27];
    res->cov[1] = (double)(S.n_edge + S.n_plane);
}
