/* TEST INFRASTRUCTURE -- CPU oracle (see vlo_oracle.h; parity unpinned for the LOAM parts).
 *
 * 6x6 float32 kernels of the Gauss-Newton step:
 *   - matX = matAtA.colPivHouseholderQr().solve(matAtB)      -> orc_solve6_colpiv_qr
 *   - SelfAdjointEigenSolver<6x6 float> at iteration 0         -> orc_eig6_jacobi
 *   - lambda_i < degenEigVal (loam_params.yaml:39,53) -> projection matP -> orc_degeneracy
 * and the reference's own D-optimality gate
 *   - gtsam_fusion/src/degerate_odometry_filter.cpp:30-46     -> orc_dopt_gate  (PINNED: in-repo)
 *
 * Eigen leaves operation order to the compiler; the orders are FROZEN here so the CUDA path can
 * be bit-exact: column-pivoted Householder QR with directly recomputed column norms, and a
 * cyclic Jacobi with the round-robin (tournament) pair order, each round's three rotations
 * computed from the same matrix, applied columns-then-rows, matrix re-symmetrised from its
 * upper triangle after every round, 8 sweeps.
 */
#include "vlo_oracle.h"
#include <math.h>
#include <string.h>
#include <float.h>

void orc_solve6_colpiv_qr(const float *Ain, const float *bin, float *x)
{
    float A[6][6], b[6];
    int perm[6];
    for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) A[i][j] = Ain[i * 6 + j]; b[i] = bin[i]; perm[i] = i; }
    /* Eigen: threshold_helper = (max col norm * eps)^2 / rows */
    float maxn2 = 0.0f;
    for (int j = 0; j < 6; j++) {
        float s = 0.0f;
        for (int i = 0; i < 6; i++) s += A[i][j] * A[i][j];
        if (s > maxn2) maxn2 = s;
    }
    float mx = sqrtf(maxn2) * FLT_EPSILON;
    float thr_helper = (mx * mx) / 6.0f;
    int nonzero = 6;
    for (int k = 0; k < 6; k++) {
        int piv = k; float best = -1.0f;
        for (int j = k; j < 6; j++) {
            float s = 0.0f;
            for (int i = k; i < 6; i++) s += A[i][j] * A[i][j];
            if (s > best) { best = s; piv = j; }
        }
        if (nonzero == 6 && best < thr_helper * (float)(6 - k)) nonzero = k;
        if (piv != k) {
            for (int i = 0; i < 6; i++) { float t = A[i][k]; A[i][k] = A[i][piv]; A[i][piv] = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        float nrm = sqrtf(best);
        if (nrm == 0.0f) continue;
        float alpha = (A[k][k] >= 0.0f) ? -nrm : nrm;
        float v[6];
        for (int i = 0; i < 6; i++) v[i] = 0.0f;
        for (int i = k; i < 6; i++) v[i] = A[i][k];
        v[k] = v[k] - alpha;
        float vn2 = 0.0f;
        for (int i = k; i < 6; i++) vn2 += v[i] * v[i];
        if (vn2 == 0.0f) continue;
        for (int j = k; j < 6; j++) {
            float dot = 0.0f;
            for (int i = k; i < 6; i++) dot += v[i] * A[i][j];
            float f = (2.0f * dot) / vn2;
            for (int i = k; i < 6; i++) A[i][j] = A[i][j] - f * v[i];
        }
        {
            float dot = 0.0f;
            for (int i = k; i < 6; i++) dot += v[i] * b[i];
            float f = (2.0f * dot) / vn2;
            for (int i = k; i < 6; i++) b[i] = b[i] - f * v[i];
        }
    }
    float y[6];
    for (int i = 0; i < 6; i++) y[i] = 0.0f;
    for (int i = nonzero - 1; i >= 0; i--) {
        float s = b[i];
        for (int j = i + 1; j < nonzero; j++) s = s - A[i][j] * y[j];
        y[i] = s / A[i][i];
    }
    for (int i = 0; i < 6; i++) x[perm[i]] = y[i];
}

static const int JROUND[5][3][2] = {
    { {0, 5}, {1, 4}, {2, 3} },
    { {0, 4}, {3, 5}, {1, 2} },
    { {0, 3}, {2, 4}, {1, 5} },
    { {0, 2}, {1, 3}, {4, 5} },
    { {0, 1}, {2, 5}, {3, 4} },
};
#define ORC_JACOBI_SWEEPS 8

void orc_eig6_jacobi(const float *Ain, float *eval, float *evec)
{
    float A[6][6], V[6][6];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { A[i][j] = Ain[i * 6 + j]; V[i][j] = (i == j) ? 1.0f : 0.0f; }
    for (int i = 0; i < 6; i++) for (int j = i + 1; j < 6; j++) A[j][i] = A[i][j];
    for (int sweep = 0; sweep < ORC_JACOBI_SWEEPS; sweep++) {
        for (int r = 0; r < 5; r++) {
            float cs[3], sn[3];
            for (int m = 0; m < 3; m++) {
                int p = JROUND[r][m][0], q = JROUND[r][m][1];
                float apq = A[p][q];
                if (apq == 0.0f) { cs[m] = 1.0f; sn[m] = 0.0f; continue; }
                float theta = (A[q][q] - A[p][p]) / (2.0f * apq);
                float t = 1.0f / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
                if (theta < 0.0f) t = -t;
                float c = 1.0f / sqrtf(t * t + 1.0f);
                cs[m] = c; sn[m] = t * c;
            }
            for (int m = 0; m < 3; m++) {           /* columns: A <- A J */
                int p = JROUND[r][m][0], q = JROUND[r][m][1];
                for (int i = 0; i < 6; i++) {
                    float aip = A[i][p], aiq = A[i][q];
                    A[i][p] = cs[m] * aip - sn[m] * aiq;
                    A[i][q] = sn[m] * aip + cs[m] * aiq;
                }
            }
            for (int m = 0; m < 3; m++) {           /* rows: A <- J^T A */
                int p = JROUND[r][m][0], q = JROUND[r][m][1];
                for (int j = 0; j < 6; j++) {
                    float apj = A[p][j], aqj = A[q][j];
                    A[p][j] = cs[m] * apj - sn[m] * aqj;
                    A[q][j] = sn[m] * apj + cs[m] * aqj;
                }
            }
            for (int m = 0; m < 3; m++) {           /* V <- V J */
                int p = JROUND[r][m][0], q = JROUND[r][m][1];
                for (int i = 0; i < 6; i++) {
                    float vip = V[i][p], viq = V[i][q];
                    V[i][p] = cs[m] * vip - sn[m] * viq;
                    V[i][q] = sn[m] * vip + cs[m] * viq;
                }
            }
            for (int i = 0; i < 6; i++) for (int j = i + 1; j < 6; j++) A[j][i] = A[i][j];
        }
    }
    int order[6] = { 0, 1, 2, 3, 4, 5 };
    for (int i = 1; i < 6; i++) {                    /* stable insertion sort ascending */
        int v = order[i]; int j = i;
        while (j >= 1 && A[v][v] < A[order[j - 1]][order[j - 1]]) { order[j] = order[j - 1]; j--; }
        order[j] = v;
    }
    for (int i = 0; i < 6; i++) {
        eval[i] = A[order[i]][order[i]];
        for (int k = 0; k < 6; k++) evec[i * 6 + k] = V[k][order[i]];
    }
}

int orc_degeneracy(const float *A, float thr, float *eval, float *P)
{
    float evec[36];
    orc_eig6_jacobi(A, eval, evec);
    int n_drop = 0;
    for (int i = 0; i < 6; i++) { if (eval[i] < thr) n_drop++; else break; }
    /* matP = V^-1 V2 with rows of V2 zeroed == sum over kept eigenvectors v v^T (V orthonormal) */
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            float s = 0.0f;
            for (int i = n_drop; i < 6; i++) s += evec[i * 6 + a] * evec[i * 6 + b];
            P[a * 6 + b] = s;
        }
    return n_drop > 0;
}

/* Eigen fixed-size 3x3 determinant (cofactor expansion along row 0), float */
static float det3(const float *H, int o)
{
#define M(r, c) H[(o + (r)) * 6 + (o + (c))]
    float a = M(0, 0) * (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1));
    float b = M(0, 1) * (M(1, 0) * M(2, 2) - M(1, 2) * M(2, 0));
    float c = M(0, 2) * (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0));
#undef M
    return (a - b) + c;
}

int orc_dopt_gate(const float *hessian36, double rot_thr, double trans_thr, float *logdet_rot, float *logdet_trans)
{
    /* degerate_odometry_filter.cpp:32-36: rotation = block<3,3>(3,3), translation = block<3,3>(0,0) */
    float rot = logf(det3(hessian36, 3));
    float trans = logf(det3(hessian36, 0));
    if (logdet_rot) *logdet_rot = rot;
    if (logdet_trans) *logdet_trans = trans;
    /* :39  NaN compares false -> message passes */
    if ((double)rot < rot_thr || (double)trans < trans_thr) return 0;
    return 1;
}
