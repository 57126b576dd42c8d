// K6  batched IMU preintegration, one warp per factor (float64).
// Replaces VILFusion::IMUManager::getFactor (gtsam_fusion/src/gtsam_fusion/IMUManager.cpp:27-74: the
// window rule -- drop samples <= t0, integrate every sample < t1 over dt = t_k - t_{k-1} (first from
// t0), then one linearly interpolated step up to t1) and the gtsam
// PreintegratedCombinedMeasurements::integrateMeasurement it calls (:50-52,64): manifold (Forster)
// update of (dR, dP, dV), bias Jacobians, and the 15x15 covariance F P F^T + G Q G^T with the noise
// set by ImuManagerRos::getImuParams (ImuManagerRos.cpp:14-36).  SURVEY.md Appendix B.
// The 3x3 algebra is computed redundantly by every lane (no divergence, no broadcasts); the two
// 15x15 products of the covariance propagation are spread over the 32 lanes through shared memory.
#include "vlo_internal.cuh"

#define IMU_WARPS 4

struct ImuParamsDev { double cov_accel, cov_gyro, cov_integration, cov_bias_acc, cov_bias_omega, cov_bias_int; };

__device__ __forceinline__ void m3mul(const double *A, const double *B, double *C)
{
    double t[9];
    #pragma unroll
    for (int i = 0; i < 3; i++)
        #pragma unroll
        for (int j = 0; j < 3; j++) t[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    #pragma unroll
    for (int i = 0; i < 9; i++) C[i] = t[i];
}
__device__ __forceinline__ void m3T(const double *A, double *At)
{
    double t[9];
    #pragma unroll
    for (int i = 0; i < 3; i++)
        #pragma unroll
        for (int j = 0; j < 3; j++) t[i * 3 + j] = A[j * 3 + i];
    #pragma unroll
    for (int i = 0; i < 9; i++) At[i] = t[i];
}
__device__ __forceinline__ void m3vec(const double *A, const double *v, double *o)
{
    double t0 = A[0] * v[0] + A[1] * v[1] + A[2] * v[2], t1 = A[3] * v[0] + A[4] * v[1] + A[5] * v[2], t2 = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
    o[0] = t0; o[1] = t1; o[2] = t2;
}
__device__ __forceinline__ void skew3(const double *v, double *S)
{
    S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
__device__ __forceinline__ void so3_expmap(const double *w, double *R, double *Jr)
{
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double W[9], WW[9];
    skew3(w, W); m3mul(W, W, WW);
    double a, b, c;
    if (th2 > 1e-20) { double th = sqrt(th2); a = sin(th) / th; b = (1.0 - cos(th)) / th2; c = (1.0 - a) / th2; }
    else { a = 1.0; b = 0.5; c = 1.0 / 6.0; }
    #pragma unroll
    for (int i = 0; i < 9; i++) {
        double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + a * W[i] + b * WW[i];
        Jr[i] = I - b * W[i] + c * WW[i];
    }
}

struct ImuState {
    double R[9], p[3], v[3], dR_dbg[9], dP_dba[9], dP_dbg[9], dV_dba[9], dV_dbg[9], dt; int n;
};

// one integrateMeasurement; cov/F/FP are this warp's shared-memory 15x15 buffers
__device__ void imu_integrate(ImuState &s, const ImuParamsDev &prm, const double *bias, const double *acc_m, const double *gyr_m,
                              double dt, double *cov, double *F, double *FP, int lane)
{
    double acc[3] = { acc_m[0] - bias[0], acc_m[1] - bias[1], acc_m[2] - bias[2] };
    double om[3] = { gyr_m[0] - bias[3], gyr_m[1] - bias[4], gyr_m[2] - bias[5] };
    double dt22 = 0.5 * dt * dt;
    double Rold[9], RoldT[9];
    #pragma unroll
    for (int i = 0; i < 9; i++) Rold[i] = s.R[i];
    m3T(Rold, RoldT);
    double b_v[3]; m3vec(RoldT, s.v, b_v);
    double xiR[3] = { dt * om[0], dt * om[1], dt * om[2] }, xiP[3], xiV[3];
    #pragma unroll
    for (int i = 0; i < 3; i++) { xiP[i] = dt * b_v[i] + dt22 * acc[i]; xiV[i] = dt * acc[i]; }
    double bRc[9], Jr[9], bRcT[9];
    so3_expmap(xiR, bRc, Jr);
    m3T(bRc, bRcT);
    m3mul(Rold, bRc, s.R);
    double dp[3], dv[3]; m3vec(Rold, xiP, dp); m3vec(Rold, xiV, dv);
    #pragma unroll
    for (int i = 0; i < 3; i++) { s.p[i] += dp[i]; s.v[i] += dv[i]; }
    s.dt += dt;

    double SxiP[9], SxiV[9], Sbv[9], tP[9], tB[9], tV[9];
    skew3(xiP, SxiP); skew3(xiV, SxiV); skew3(b_v, Sbv);
    m3mul(bRcT, SxiP, tP); m3mul(bRcT, Sbv, tB); m3mul(bRcT, SxiV, tV);
    double theta_H_bg[9], vel_H_ba[9];
    #pragma unroll
    for (int i = 0; i < 9; i++) { theta_H_bg[i] = -Jr[i] * dt; vel_H_ba[i] = -bRcT[i] * dt; }

    // bias Jacobians
    double Sacc[9], D_acc_R[9], D_acc_bg[9], incrRt[9], newdR[9];
    skew3(acc, Sacc); m3mul(Rold, Sacc, D_acc_R);
    #pragma unroll
    for (int i = 0; i < 9; i++) D_acc_R[i] = -D_acc_R[i];
    m3mul(D_acc_R, s.dR_dbg, D_acc_bg);
    m3T(bRc, incrRt);
    m3mul(incrRt, s.dR_dbg, newdR);
    #pragma unroll
    for (int i = 0; i < 9; i++) newdR[i] -= Jr[i] * dt;
    #pragma unroll
    for (int i = 0; i < 9; i++) {
        s.dP_dba[i] += s.dV_dba[i] * dt - dt22 * Rold[i];
        s.dP_dbg[i] += dt * s.dV_dbg[i] + dt22 * D_acc_bg[i];
        s.dV_dba[i] += -Rold[i] * dt;
        s.dV_dbg[i] += D_acc_bg[i] * dt;
        s.dR_dbg[i] = newdR[i];
    }
    s.n++;

    // F (15x15) into shared memory, lane-strided fill
    for (int k = lane; k < 225; k += 32) {
        int i = k / 15, j = k % 15;
        double v = 0.0;
        int bi = i / 3, bj = j / 3, ii = i % 3, jj = j % 3, e = ii * 3 + jj;
        if (bi == 0) { if (bj == 0) v = bRcT[e]; else if (bj == 4) v = theta_H_bg[e]; }
        else if (bi == 1) { if (bj == 0) v = -tP[e] + dt * tB[e]; else if (bj == 1) v = bRcT[e]; else if (bj == 2) v = dt * bRcT[e]; }
        else if (bi == 2) { if (bj == 0) v = -tV[e]; else if (bj == 2) v = bRcT[e]; else if (bj == 3) v = vel_H_ba[e]; }
        else if (i == j) v = 1.0;
        F[k] = v;
    }
    __syncwarp();
    for (int k = lane; k < 225; k += 32) {
        int i = k / 15, j = k % 15;
        double acc2 = 0.0;
        for (int q = 0; q < 15; q++) acc2 += F[i * 15 + q] * cov[q * 15 + j];
        FP[k] = acc2;
    }
    __syncwarp();
    double aC = prm.cov_accel + prm.cov_bias_int, wC = prm.cov_gyro + prm.cov_bias_int;
    double vv[9], rr[9], t3[9];
    m3T(vel_H_ba, t3); m3mul(vel_H_ba, t3, vv);
    m3T(theta_H_bg, t3); m3mul(theta_H_bg, t3, rr);
    for (int k = lane; k < 225; k += 32) {
        int i = k / 15, j = k % 15;
        double acc2 = 0.0;
        for (int q = 0; q < 15; q++) acc2 += FP[i * 15 + q] * F[j * 15 + q];
        int bi = i / 3, bj = j / 3, e = (i % 3) * 3 + (j % 3);
        double g = 0.0;
        if (bi == bj) {
            if (bi == 0) g = (1.0 / dt) * wC * rr[e];
            else if (bi == 2) g = (1.0 / dt) * aC * vv[e];
            else if (i == j) g = dt * (bi == 1 ? prm.cov_integration : (bi == 3 ? prm.cov_bias_acc : prm.cov_bias_omega));
        }
        cov[k] = acc2 + g;     // each (i,j) is owned by one lane; FP and F are read-only here
    }
    __syncwarp();
}

__global__ void __launch_bounds__(IMU_WARPS * 32) k6_imu_preintegrate(ImuParamsDev prm, const double *t, const double *acc, const double *gyro,
                                                                       int n, const double *t0s, const double *t1s, const double *bias6,
                                                                       int n_factors, vlo_preint *out)
{
    __shared__ double s_cov[IMU_WARPS][225], s_F[IMU_WARPS][225], s_FP[IMU_WARPS][225];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * IMU_WARPS + warp;
    if (f >= n_factors) return;
    double *cov = s_cov[warp], *F = s_F[warp], *FP = s_FP[warp];
    for (int k = lane; k < 225; k += 32) cov[k] = 0.0;
    __syncwarp();
    const double t0 = t0s[f], t1 = t1s[f];
    double bias[6];
    #pragma unroll
    for (int i = 0; i < 6; i++) bias[i] = bias6[i];
    ImuState s;
    #pragma unroll
    for (int i = 0; i < 9; i++) { s.R[i] = (i % 4 == 0) ? 1.0 : 0.0; s.dR_dbg[i] = 0; s.dP_dba[i] = 0; s.dP_dbg[i] = 0; s.dV_dba[i] = 0; s.dV_dbg[i] = 0; }
    #pragma unroll
    for (int i = 0; i < 3; i++) { s.p[i] = 0; s.v[i] = 0; }
    s.dt = 0; s.n = 0;
    // first sample with t > t0 (IMUManager.cpp:33-39 drops everything <= startTime)
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (t[mid] <= t0) lo = mid + 1; else hi = mid; }
    int k = lo;
    double prev_a[3] = { 0, 0, 0 }, prev_g[3] = { 0, 0, 0 };
    if (k > 0) {
        #pragma unroll
        for (int i = 0; i < 3; i++) { prev_a[i] = acc[3 * (k - 1) + i]; prev_g[i] = gyro[3 * (k - 1) + i]; }
    }
    double prev_t = t0;
    while (k < n && t[k] < t1) {                                        // :44-52
        double a[3] = { acc[3 * k], acc[3 * k + 1], acc[3 * k + 2] }, g[3] = { gyro[3 * k], gyro[3 * k + 1], gyro[3 * k + 2] };
        imu_integrate(s, prm, bias, a, g, t[k] - prev_t, cov, F, FP, lane);
        prev_t = t[k];
        #pragma unroll
        for (int i = 0; i < 3; i++) { prev_a[i] = a[i]; prev_g[i] = g[i]; }
        k++;
    }
    if (k < n) {                                                        // :55-66
        double fct = (t1 - prev_t) / (t[k] - prev_t);
        double ia[3], ig[3];
        #pragma unroll
        for (int i = 0; i < 3; i++) {
            ia[i] = (fct * acc[3 * k + i]) + ((1.0 - fct) * prev_a[i]);
            ig[i] = (fct * gyro[3 * k + i]) + ((1.0 - fct) * prev_g[i]);
        }
        imu_integrate(s, prm, bias, ia, ig, t1 - prev_t, cov, F, FP, lane);
    }
    vlo_preint *o = out + f;
    if (lane == 0) {
        for (int i = 0; i < 9; i++) { o->dR[i] = s.R[i]; o->dR_dbg[i] = s.dR_dbg[i]; o->dP_dba[i] = s.dP_dba[i]; o->dP_dbg[i] = s.dP_dbg[i]; o->dV_dba[i] = s.dV_dba[i]; o->dV_dbg[i] = s.dV_dbg[i]; }
        for (int i = 0; i < 3; i++) { o->dP[i] = s.p[i]; o->dV[i] = s.v[i]; }
        o->dt = s.dt; o->n_integrated = s.n; o->_pad = 0;
    }
    for (int q = lane; q < 225; q += 32) o->cov[q] = cov[q];
}

int vlo_launch_imu(vlo_handle *h, const double *d_t, const double *d_acc, const double *d_gyro, int n_samples,
                   const double *d_t0, const double *d_t1, const double *d_bias, int n_factors, vlo_preint *d_out)
{
    ImuParamsDev prm = { h->cfg.cov_accel, h->cfg.cov_gyro, h->cfg.cov_integration, h->cfg.cov_bias_acc, h->cfg.cov_bias_omega,
                         h->cfg.cov_bias_acc_omega_int };
    int blocks = (n_factors + IMU_WARPS - 1) / IMU_WARPS;
    VLO_PROF(h, ST_IMU, (k6_imu_preintegrate<<<blocks, IMU_WARPS * 32, 0, h->stream>>>(prm, d_t, d_acc, d_gyro, n_samples, d_t0, d_t1, d_bias, n_factors, d_out)));
    h->launches += 1;
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}
