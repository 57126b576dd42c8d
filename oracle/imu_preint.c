/* TEST INFRASTRUCTURE -- CPU oracle (see vlo_oracle.h).
 *
 * (1) PINNED, in-repo: the windowing / interpolation rule of
 *     VILFusion::IMUManager::getFactor, gtsam_fusion/src/gtsam_fusion/IMUManager.cpp:27-74,
 *     with the noise parameters of ImuManagerRos::getImuParams (ImuManagerRos.cpp:14-36,
 *     fusion_params.yaml:20-27) and the known-answer test gtsam_fusion/test/UnitTests.cpp:30-66.
 * (2) UNPINNED, third-party: what IMUManager.cpp:50-52,64 calls --
 *     gtsam::PreintegratedCombinedMeasurements::integrateMeasurement (GTSAM 4.0.3..4.2 by API use,
 *     gtsam_fusion/CMakeLists.txt:15 has no version).  Restated from the published algorithm
 *     (Forster et al., "On-Manifold Preintegration", ManifoldPreintegration::update +
 *     NavState::update + the 15x15 covariance propagation of CombinedImuFactor.cpp, 4.0.x form).
 *     State order of the covariance: (theta, p, v, bias_acc, bias_gyro).  float64.
 */
#include "vlo_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

typedef double m3[9];

static void m3_mul(const double *A, const double *B, double *C)
{
    double t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += A[i * 3 + k] * B[k * 3 + j];
        t[i * 3 + j] = s;
    }
    memcpy(C, t, sizeof(t));
}
static void m3_T(const double *A, double *At) { double t[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[i * 3 + j] = A[j * 3 + i]; memcpy(At, t, sizeof(t)); }
static void m3_vec(const double *A, const double *v, double *o) { double t[3]; for (int i = 0; i < 3; i++) t[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2]; memcpy(o, t, sizeof(t)); }
static void skew(const double *v, double *S) { S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0; }

/* SO(3) exponential and its right Jacobian (gtsam::so3::ExpmapFunctor / DexpFunctor) */
static void so3_expmap(const double *w, double *R, double *Jr)
{
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double W[9], WW[9];
    skew(w, W);
    m3_mul(W, W, WW);
    double a, b, c;   /* R = I + a W + b WW ; Jr = I - b W + c WW */
    if (th2 > 1e-20) {   /* near-zero: first-order (gtsam nearZero threshold on theta^2 <= eps) */
        double th = sqrt(th2);
        a = sin(th) / th;
        b = (1.0 - cos(th)) / th2;
        c = (1.0 - a) / th2;
    } else {
        a = 1.0; b = 0.5; c = 1.0 / 6.0;
    }
    for (int i = 0; i < 9; i++) {
        double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + a * W[i] + b * WW[i];
        if (Jr) Jr[i] = I - b * W[i] + c * WW[i];
    }
}

typedef struct {
    double R[9], p[3], v[3];
    double dR_dbg[9], dP_dba[9], dP_dbg[9], dV_dba[9], dV_dbg[9];
    double cov[225];
    double dt;
    int n;
} pim_state;

static void pim_reset(pim_state *s)
{
    memset(s, 0, sizeof(*s));
    s->R[0] = s->R[4] = s->R[8] = 1.0;
}

#define C15(M, i, j) (M)[(i) * 15 + (j)]

static void pim_integrate(pim_state *s, const orc_imu_params *prm, const double *bias, const double *acc_m, const double *gyr_m, double dt)
{
    double acc[3] = { acc_m[0] - bias[0], acc_m[1] - bias[1], acc_m[2] - bias[2] };
    double om[3] = { gyr_m[0] - bias[3], gyr_m[1] - bias[4], gyr_m[2] - bias[5] };
    double dt22 = 0.5 * dt * dt;
    double Rold[9]; memcpy(Rold, s->R, sizeof(Rold));
    double RoldT[9]; m3_T(Rold, RoldT);

    /* NavState::update */
    double b_v[3]; m3_vec(RoldT, s->v, b_v);                 /* bodyVelocity */
    double xiR[3] = { dt * om[0], dt * om[1], dt * om[2] };
    double xiP[3], xiV[3];
    for (int i = 0; i < 3; i++) { xiP[i] = dt * b_v[i] + dt22 * acc[i]; xiV[i] = dt * acc[i]; }
    double bRc[9], Jr[9];
    so3_expmap(xiR, bRc, Jr);
    double bRcT[9]; m3_T(bRc, bRcT);
    double Rnew[9]; m3_mul(Rold, bRc, Rnew);
    double dp[3], dv[3]; m3_vec(Rold, xiP, dp); m3_vec(Rold, xiV, dv);
    for (int i = 0; i < 3; i++) { s->p[i] += dp[i]; s->v[i] += dv[i]; }
    memcpy(s->R, Rnew, sizeof(Rnew));
    s->dt += dt;

    /* A (9x9), B (9x3), C (9x3) */
    double A[81]; memset(A, 0, sizeof(A));
    double SxiP[9], SxiV[9], Sbv[9], tmp[9], tmp2[9];
    skew(xiP, SxiP); skew(xiV, SxiV); skew(b_v, Sbv);
    /* rows R */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[i * 9 + j] = bRcT[i * 3 + j];
    /* rows P: [-bRc^T [xiP]x + dt bRc^T [b_v]x , bRc^T , dt bRc^T] */
    m3_mul(bRcT, SxiP, tmp); m3_mul(bRcT, Sbv, tmp2);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        A[(3 + i) * 9 + j] = -tmp[i * 3 + j] + dt * tmp2[i * 3 + j];
        A[(3 + i) * 9 + 3 + j] = bRcT[i * 3 + j];
        A[(3 + i) * 9 + 6 + j] = dt * bRcT[i * 3 + j];
    }
    /* rows V: [-bRc^T [xiV]x , 0 , bRc^T] */
    m3_mul(bRcT, SxiV, tmp);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        A[(6 + i) * 9 + j] = -tmp[i * 3 + j];
        A[(6 + i) * 9 + 6 + j] = bRcT[i * 3 + j];
    }
    /* B = [0; bRc^T dt22; bRc^T dt], C = [Jr dt; 0; 0] */
    double theta_H_bg[9], vel_H_ba[9];
    for (int i = 0; i < 9; i++) { theta_H_bg[i] = -Jr[i] * dt; vel_H_ba[i] = -bRcT[i] * dt; }

    /* bias Jacobians (ManifoldPreintegration::update) */
    double D_acc_R[9], Sacc[9];
    skew(acc, Sacc);
    m3_mul(Rold, Sacc, D_acc_R);
    for (int i = 0; i < 9; i++) D_acc_R[i] = -D_acc_R[i];       /* d(R a)/dR = -R [a]x */
    double D_acc_bg[9]; m3_mul(D_acc_R, s->dR_dbg, D_acc_bg);
    double incrR[9], Dincr[9];
    so3_expmap(xiR, incrR, Dincr);
    double incrRt[9]; m3_T(incrR, incrRt);
    double newdRdbg[9]; m3_mul(incrRt, s->dR_dbg, newdRdbg);
    for (int i = 0; i < 9; i++) newdRdbg[i] -= Dincr[i] * dt;
    for (int i = 0; i < 9; i++) {
        s->dP_dba[i] += s->dV_dba[i] * dt - dt22 * Rold[i];
        s->dP_dbg[i] += dt * s->dV_dbg[i] + dt22 * D_acc_bg[i];
        s->dV_dba[i] += -Rold[i] * dt;
        s->dV_dbg[i] += D_acc_bg[i] * dt;
    }
    memcpy(s->dR_dbg, newdRdbg, sizeof(newdRdbg));

    /* covariance: F P F^T + G Q G^T (CombinedImuFactor.cpp, 4.0.x) */
    double F[225]; memset(F, 0, sizeof(F));
    for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) C15(F, i, j) = A[i * 9 + j];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        C15(F, i, 12 + j) = theta_H_bg[i * 3 + j];
        C15(F, 6 + i, 9 + j) = vel_H_ba[i * 3 + j];
    }
    for (int i = 9; i < 15; i++) C15(F, i, i) = 1.0;
    double G[225]; memset(G, 0, sizeof(G));
    double aC = prm->cov_accel + prm->cov_bias_acc_omega_int;     /* aCov + biasAccOmegaInt.block(0,0) */
    double wC = prm->cov_gyro + prm->cov_bias_acc_omega_int;      /* wCov + biasAccOmegaInt.block(3,3) */
    double vv[9], rr[9], t3[9];
    m3_T(vel_H_ba, t3); m3_mul(vel_H_ba, t3, vv);
    m3_T(theta_H_bg, t3); m3_mul(theta_H_bg, t3, rr);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        C15(G, i, j) = (1.0 / dt) * wC * rr[i * 3 + j];
        C15(G, 6 + i, 6 + j) = (1.0 / dt) * aC * vv[i * 3 + j];
    }
    for (int i = 0; i < 3; i++) {
        C15(G, 3 + i, 3 + i) = dt * prm->cov_integration;
        C15(G, 9 + i, 9 + i) = dt * prm->cov_bias_acc;
        C15(G, 12 + i, 12 + i) = dt * prm->cov_bias_omega;
    }
    /* off-diagonal block uses biasAccOmegaInt.block(3,0), zero for the reference's isotropic setting */
    double FP[225], Pn[225];
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) {
        double sacc = 0; for (int k = 0; k < 15; k++) sacc += C15(F, i, k) * C15(s->cov, k, j);
        C15(FP, i, j) = sacc;
    }
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) {
        double sacc = 0; for (int k = 0; k < 15; k++) sacc += C15(FP, i, k) * C15(F, j, k);
        C15(Pn, i, j) = sacc + C15(G, i, j);
    }
    memcpy(s->cov, Pn, sizeof(Pn));
    s->n++;
}

void orc_imu_get_factor(const orc_imu_params *prm, const double *t, const double *acc, const double *gyro, int n,
                        double t0, double t1, const double *bias6, orc_preint *out)
{
    /* IMUManager.cpp:33-39: drop everything with time <= startTime, remembering the last dropped */
    pim_state s; pim_reset(&s);
    int k = 0;
    double prev_a[3] = { 0, 0, 0 }, prev_g[3] = { 0, 0, 0 };   /* oracle-defined when nothing was dropped */
    while (k < n && t[k] <= t0) { memcpy(prev_a, acc + 3 * k, sizeof(prev_a)); memcpy(prev_g, gyro + 3 * k, sizeof(prev_g)); k++; }
    double prev_t = t0;                                           /* :42 */
    /* :44-52 */
    while (k < n && t[k] < t1) {
        pim_integrate(&s, prm, bias6, acc + 3 * k, gyro + 3 * k, t[k] - prev_t);
        prev_t = t[k]; memcpy(prev_a, acc + 3 * k, sizeof(prev_a)); memcpy(prev_g, gyro + 3 * k, sizeof(prev_g));
        k++;
    }
    /* :55-66 final interpolated step (the sample stays in the buffer) */
    if (k < n) {
        double f = (t1 - prev_t) / (t[k] - prev_t);
        double ia[3], ig[3];
        for (int i = 0; i < 3; i++) {
            ia[i] = (f * acc[3 * k + i]) + ((1.0 - f) * prev_a[i]);
            ig[i] = (f * gyro[3 * k + i]) + ((1.0 - f) * prev_g[i]);
        }
        pim_integrate(&s, prm, bias6, ia, ig, t1 - prev_t);
    }
    memcpy(out->dR, s.R, sizeof(s.R)); memcpy(out->dP, s.p, sizeof(s.p)); memcpy(out->dV, s.v, sizeof(s.v));
    memcpy(out->dR_dbg, s.dR_dbg, sizeof(s.dR_dbg)); memcpy(out->dP_dba, s.dP_dba, sizeof(s.dP_dba));
    memcpy(out->dP_dbg, s.dP_dbg, sizeof(s.dP_dbg)); memcpy(out->dV_dba, s.dV_dba, sizeof(s.dV_dba));
    memcpy(out->dV_dbg, s.dV_dbg, sizeof(s.dV_dbg)); memcpy(out->cov, s.cov, sizeof(s.cov));
    out->dt = s.dt; out->n_integrated = s.n;
}

typedef struct {
    const orc_imu_params *p; const double *t, *acc, *gyro; int n;
    const double *t0, *t1, *bias6; orc_preint *out; int f_begin, f_end;
} imu_job;

static void *imu_worker(void *arg)
{
    imu_job *j = (imu_job *)arg;
    for (int f = j->f_begin; f < j->f_end; f++) {
        /* window lookup by binary search; semantics identical to a fresh IMUManager buffer
         * holding every sample (the last sample <= t0 is kept as the "dropped" one) */
        int lo = 0, hi = j->n;
        while (lo < hi) { int mid = (lo + hi) / 2; if (j->t[mid] <= j->t0[f]) lo = mid + 1; else hi = mid; }
        int begin = lo > 0 ? lo - 1 : 0;
        int end = begin;
        while (end < j->n && j->t[end] < j->t1[f]) end++;
        if (end < j->n) end++;
        orc_imu_get_factor(j->p, j->t + begin, j->acc + 3 * begin, j->gyro + 3 * begin, end - begin,
                           j->t0[f], j->t1[f], j->bias6, &j->out[f]);
    }
    return NULL;
}

void orc_imu_batch(const orc_imu_params *p, const double *t, const double *acc, const double *gyro, int n,
                   const double *t0, const double *t1, const double *bias6, int n_factors, orc_preint *out,
                   int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; imu_job jobs[256];
    for (int k = 0; k < n_threads; k++) {
        jobs[k] = (imu_job){ p, t, acc, gyro, n, t0, t1, bias6, out,
                             (int)((long long)n_factors * k / n_threads), (int)((long long)n_factors * (k + 1) / n_threads) };
        if (n_threads == 1) imu_worker(&jobs[k]); else pthread_create(&th[k], NULL, imu_worker, &jobs[k]);
    }
    if (n_threads > 1) for (int k = 0; k < n_threads; k++) pthread_join(th[k], NULL);
}

/* SensorManagerRos::poseDiff, SensorManagerRos.cpp:122-158.  pose = x y z qw qx qy qz */
static void quat_mul(const double *a, const double *b, double *o)
{
    double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
static void quat_inv(const double *q, double *o)
{
    /* Eigen::Quaternion::inverse(): conjugate / squaredNorm */
    double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    o[0] = q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = -q[3] / n2;
}
static void quat_rot(const double *q, const double *v, double *o)
{
    /* Eigen quaternion * vector: v + 2 w (u x v) + 2 u x (u x v) */
    double u[3] = { q[1], q[2], q[3] };
    double uv[3] = { u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0] };
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    double uuv[3] = { u[1] * uv[2] - u[2] * uv[1], u[2] * uv[0] - u[0] * uv[2], u[0] * uv[1] - u[1] * uv[0] };
    for (int i = 0; i < 3; i++) o[i] = v[i] + q[0] * uv[i] + uuv[i];
}
void orc_pose_diff(const double *before7, const double *after7, double *out7)
{
    double dx[3] = { after7[0] - before7[0], after7[1] - before7[1], after7[2] - before7[2] };   /* :142 */
    double q1i[4]; quat_inv(before7 + 3, q1i);
    quat_rot(q1i, dx, out7);                                                                      /* :143 */
    quat_mul(after7 + 3, q1i, out7 + 3);                                                          /* :148 q2 * q1^-1 */
}
