// Per-point linearisation of the scan-to-map step (LaserMapping; oracle/laser_mapping.c is the frozen
// operation order): pointAssociateToMap, the 3x3 covariance eigen-decomposition of the corner test, the 5x3 least-squares
// plane of the surface test and the two coefficient rules.  A header of its own so that the kernels (k5_mapping.cu) and the
// host emulation harness (tests/host/) compile the very same text.
#pragma once
#include "vlo_internal.cuh"
#include <cfloat>

__device__ __forceinline__ float4 to_map(const float *T, const float *trig, float4 pi)
{
    float x = pi.x, y = pi.y, z = pi.z;
    float sx = trig[0], cx = trig[1], sy = trig[2], cy = trig[3], sz = trig[4], cz = trig[5];
    float x0 = x; x = cz * x0 - sz * y; y = sz * x0 + cz * y;
    float y0 = y; y = cx * y0 - sx * z; z = sx * y0 + cx * z;
    x0 = x;       x = cy * x0 + sy * z; z = cy * z - sy * x0;
    return make_float4(x + T[3], y + T[4], z + T[5], pi.w);
}

// cyclic Jacobi on a symmetric 3x3; eval ascending, evec[k*3+i] = component i of eigenvector k
__device__ inline void eig3_jacobi(const float *Ain, float *eval, float *evec)
{
    float A[3][3], V[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { A[i][j] = Ain[i * 3 + j]; V[i][j] = (i == j) ? 1.0f : 0.0f; }
    for (int sweep = 0; sweep < 6; sweep++) {
        #pragma unroll
        for (int m = 0; m < 3; m++) {
            const int p = (m == 2) ? 1 : 0, q = (m == 0) ? 1 : 2;
            float apq = A[p][q];
            if (apq == 0.0f) continue;
            float theta = (A[q][q] - A[p][p]) / (2.0f * apq);
            float t = 1.0f / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
            if (theta < 0.0f) t = -t;
            float c = 1.0f / sqrtf(t * t + 1.0f), s = t * c;
            #pragma unroll
            for (int i = 0; i < 3; i++) { float aip = A[i][p], aiq = A[i][q]; A[i][p] = c * aip - s * aiq; A[i][q] = s * aip + c * aiq; }
            #pragma unroll
            for (int j = 0; j < 3; j++) { float apj = A[p][j], aqj = A[q][j]; A[p][j] = c * apj - s * aqj; A[q][j] = s * apj + c * aqj; }
            #pragma unroll
            for (int i = 0; i < 3; i++) { float vip = V[i][p], viq = V[i][q]; V[i][p] = c * vip - s * viq; V[i][q] = s * vip + c * viq; }
            A[q][p] = A[p][q];
            const int r = 3 - p - q;
            A[r][p] = A[p][r]; A[r][q] = A[q][r];
        }
    }
    // stable insertion sort of the three eigenpairs by eigenvalue (strict <), written as adjacent conditional swaps on
    // named registers: no run-time array index, so nothing of this function lives in local memory
    float e0 = A[0][0], e1 = A[1][1], e2 = A[2][2];
    float c0[3] = { V[0][0], V[1][0], V[2][0] }, c1[3] = { V[0][1], V[1][1], V[2][1] }, c2[3] = { V[0][2], V[1][2], V[2][2] };
    #define EIG3_SWAP(ea, ca, eb, cb) { float t_ = ea; ea = eb; eb = t_; _Pragma("unroll") for (int i_ = 0; i_ < 3; i_++) { float u_ = ca[i_]; ca[i_] = cb[i_]; cb[i_] = u_; } }
    if (e1 < e0) EIG3_SWAP(e0, c0, e1, c1)
    if (e2 < e1) { EIG3_SWAP(e1, c1, e2, c2) if (e1 < e0) EIG3_SWAP(e0, c0, e1, c1) }
    #undef EIG3_SWAP
    eval[0] = e0; eval[1] = e1; eval[2] = e2;
    #pragma unroll
    for (int i = 0; i < 3; i++) { evec[i] = c0[i]; evec[3 + i] = c1[i]; evec[6 + i] = c2[i]; }
}

// One Householder step of the 5x3 column-pivoted QR with every index a compile-time constant (K = step).
template <int K>
__device__ __forceinline__ void lstsq53_step(float (&A)[5][3], float (&b)[5], int (&perm)[3], int &nonzero, float thr_helper)
{
    int piv = K; float best = -1.0f;
    #pragma unroll
    for (int j = K; j < 3; j++) {
        float s = 0.0f;
        #pragma unroll
        for (int i = K; i < 5; i++) s += A[i][j] * A[i][j];
        if (s > best) { best = s; piv = j; }
    }
    if (nonzero == 3 && best < thr_helper * (float)(5 - K)) nonzero = K;
    #pragma unroll
    for (int j = K + 1; j < 3; j++)
        if (piv == j) {
            #pragma unroll
            for (int i = 0; i < 5; i++) { float t = A[i][K]; A[i][K] = A[i][j]; A[i][j] = t; }
            int t = perm[K]; perm[K] = perm[j]; perm[j] = t;
        }
    const float nrm = sqrtf(best);
    if (nrm == 0.0f) return;
    const float alpha = (A[K][K] >= 0.0f) ? -nrm : nrm;
    float v[5];
    #pragma unroll
    for (int i = K; i < 5; i++) v[i] = A[i][K];
    v[K] = v[K] - alpha;
    float vn2 = 0.0f;
    #pragma unroll
    for (int i = K; i < 5; i++) vn2 += v[i] * v[i];
    if (vn2 == 0.0f) return;
    #pragma unroll
    for (int j = K; j < 3; j++) {
        float dot = 0.0f;
        #pragma unroll
        for (int i = K; i < 5; i++) dot += v[i] * A[i][j];
        const float f = (2.0f * dot) / vn2;
        #pragma unroll
        for (int i = K; i < 5; i++) A[i][j] = A[i][j] - f * v[i];
    }
    float dot = 0.0f;
    #pragma unroll
    for (int i = K; i < 5; i++) dot += v[i] * b[i];
    const float f = (2.0f * dot) / vn2;
    #pragma unroll
    for (int i = K; i < 5; i++) b[i] = b[i] - f * v[i];
}

// min ||A x + 1|| for the 5 x 3 neighbour matrix: column-pivoted Householder QR (Q1 of the oracle, orc_lstsq53), same
// operations in the same order; the pivot swaps, the rank-dependent back-substitution and the final permutation are
// spelled out on constant indices so that A, b, v stay in registers (the run-time-indexed version kept them in local
// memory: 623 LDL/STL in the kernel)
__device__ __forceinline__ void lstsq53(const float (*Ain)[3], float *x)
{
    float A[5][3], b[5];
    int perm[3] = { 0, 1, 2 };
    #pragma unroll
    for (int i = 0; i < 5; i++) {
        #pragma unroll
        for (int j = 0; j < 3; j++) A[i][j] = Ain[i][j];
        b[i] = -1.0f;
    }
    float maxn2 = 0.0f;
    #pragma unroll
    for (int j = 0; j < 3; j++) {
        float s = 0.0f;
        #pragma unroll
        for (int i = 0; i < 5; i++) s += A[i][j] * A[i][j];
        if (s > maxn2) maxn2 = s;
    }
    const float mx = sqrtf(maxn2) * FLT_EPSILON;
    const float thr_helper = (mx * mx) / 5.0f;
    int nonzero = 3;
    lstsq53_step<0>(A, b, perm, nonzero, thr_helper);
    lstsq53_step<1>(A, b, perm, nonzero, thr_helper);
    lstsq53_step<2>(A, b, perm, nonzero, thr_helper);
    float y0 = 0.f, y1 = 0.f, y2 = 0.f;
    if (nonzero == 3) {
        y2 = b[2] / A[2][2];
        y1 = (b[1] - A[1][2] * y2) / A[1][1];
        float s0 = b[0]; s0 = s0 - A[0][1] * y1; s0 = s0 - A[0][2] * y2;
        y0 = s0 / A[0][0];
    } else if (nonzero == 2) {
        y1 = b[1] / A[1][1];
        y0 = (b[0] - A[0][1] * y1) / A[0][0];
    } else if (nonzero == 1) {
        y0 = b[0] / A[0][0];
    }
    #pragma unroll
    for (int c = 0; c < 3; c++) x[c] = (perm[0] == c) ? y0 : ((perm[1] == c) ? y1 : y2);
}

__device__ inline bool map_edge_coeff(float4 sel, const float4 *nb, float *coeff)
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
    eig3_jacobi(M, ev, evec);
    if (!(ev[2] > 3.0f * ev[1])) return false;
    float x0 = sel.x, y0 = sel.y, z0 = sel.z;
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
    return (double)s > 0.1;
}

__device__ inline bool map_plane_coeff(float4 sel, const float4 *nb, float *coeff)
{
    float A0[5][3], X0[3];
    for (int j = 0; j < 5; j++) { A0[j][0] = nb[j].x; A0[j][1] = nb[j].y; A0[j][2] = nb[j].z; }
    lstsq53(A0, X0);
    float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1.0f;
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa = pa / ps; pb = pb / ps; pc = pc / ps; pd = pd / ps;
    for (int j = 0; j < 5; j++)
        if ((double)fabsf(pa * nb[j].x + pb * nb[j].y + pc * nb[j].z + pd) > 0.2) return false;
    float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
    float s = 1.0f - 0.9f * fabsf(pd2) / sqrtf(sqrtf(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z));
    coeff[0] = s * pa; coeff[1] = s * pb; coeff[2] = s * pc; coeff[3] = s * pd2;
    return (double)s > 0.1;
}

