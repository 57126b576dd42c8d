// C-ABI, part 2: scan-to-map, IMU batch, online tick and the small host-side helpers
// (poseDiff, D-optimality gate on a host matrix, pose accumulation).
#include "vlo_internal.cuh"
#include <cstring>
#include <cmath>
#include <algorithm>
#include <vector>

static int ensure_pinned2(vlo_handle *h, size_t bytes)
{
    if (bytes <= h->pinned_bytes) return VLO_OK;
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr; h->pinned_bytes = 0;
    VLO_CUDA(cudaMallocHost(&h->pinned, bytes));
    h->pinned_bytes = bytes;
    return VLO_OK;
}

// ---------------------------------------------------------------------------------------------
// covariance of a registration result (R4: sigma^2 (AtA)^-1, float64), finished on the host: the
// device staged sum((s d)^2) in cov[0] and the correspondence count in cov[1]
static bool inv6d_host(const double *A, double *Ai)
{
    double M[6][12];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { M[i][j] = A[i * 6 + j]; M[i][6 + j] = (i == j) ? 1.0 : 0.0; }
    for (int k = 0; k < 6; k++) {
        int p = k; double mx = fabs(M[k][k]);
        for (int i = k + 1; i < 6; i++) if (fabs(M[i][k]) > mx) { mx = fabs(M[i][k]); p = i; }
        if (mx == 0.0) return false;
        if (p != k) for (int j = 0; j < 12; j++) { double t = M[k][j]; M[k][j] = M[p][j]; M[p][j] = t; }
        double d = M[k][k];
        for (int j = 0; j < 12; j++) M[k][j] /= d;
        for (int i = 0; i < 6; i++) if (i != k) { double f = M[i][k]; for (int j = 0; j < 12; j++) M[i][j] -= f * M[k][j]; }
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Ai[i * 6 + j] = M[i][6 + j];
    return true;
}

static float det3h(const float *H, int off);

// vlo_config.hessian_order = 1: the published matrices leave in (tx ty tz rx ry rz) order -- block(0,0) is then really
// the translation block the filter labels it (degerate_odometry_filter.cpp:32-33) -- and the gate is
// re-evaluated on the permuted matrix exactly as the filter would read it
static void apply_hessian_order(vlo_result *r, const vlo_config *cfg)
{
    if (!cfg || cfg->hessian_order != 1) return;
    float Hp[36]; double Cp[36];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
        const int si = (i + 3) % 6, sj = (j + 3) % 6;
        Hp[i * 6 + j] = r->hessian[si * 6 + sj]; Cp[i * 6 + j] = r->cov[si * 6 + sj];
    }
    memcpy(r->hessian, Hp, sizeof(Hp)); memcpy(r->cov, Cp, sizeof(Cp));
    if (r->n_corr_edge + r->n_corr_plane > 0) {
        const float rot = logf(det3h(r->hessian, 3)), trans = logf(det3h(r->hessian, 0));
        r->logdet_rot = rot; r->logdet_trans = trans;
        r->pass_dopt = ((double)rot < (double)cfg->dopt_rot_threshold || (double)trans < (double)cfg->dopt_trans_threshold) ? 0 : 1;
    }
}

void vlo_finish_cov_host(vlo_result *r, const vlo_config *cfg)
{
    int n = r->n_corr_edge + r->n_corr_plane;
    if (n <= 0) { for (int i = 0; i < 36; i++) r->cov[i] = 0.0; return; }
    double ssq = r->cov[0];
    double dof = n > 6 ? (double)(n - 6) : 1.0, sigma2 = ssq / dof;
    double Hd[36], Hi[36];
    for (int i = 0; i < 36; i++) Hd[i] = (double)r->hessian[i];
    if (inv6d_host(Hd, Hi)) for (int i = 0; i < 36; i++) r->cov[i] = sigma2 * Hi[i];
    else for (int i = 0; i < 36; i++) r->cov[i] = NAN;
    apply_hessian_order(r, cfg);
}

// ---------------------------------------------------------------------------------------------
// scan-to-map
extern "C" int vlo_map_build(vlo_handle *h, const float *corner, int n_corner, const float *surf, int n_surf, int on_device)
{
    if (!h || n_corner < 0 || n_surf < 0 || (n_corner > 0 && !corner) || (n_surf > 0 && !surf)) return VLO_ERR_INVALID_ARG;
    if (h->cfg.max_map_points <= 0) { h->err = "handle created with max_map_points = 0"; return VLO_ERR_STATE; }
    if (n_corner > h->cfg.max_map_points || n_surf > h->cfg.max_map_points) { h->err = "map exceeds max_map_points"; return VLO_ERR_CAPACITY; }
    cudaSetDevice(h->cfg.device);
    cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (n_corner) VLO_CUDA(cudaMemcpyAsync(h->map_pts[0], corner, sizeof(float4) * (size_t)n_corner, kind, h->stream));
    if (n_surf) VLO_CUDA(cudaMemcpyAsync(h->map_pts[1], surf, sizeof(float4) * (size_t)n_surf, kind, h->stream));
    int mn[8] = { 0, 0, n_corner, 0, n_surf, 0, 0, 0 };
    VLO_CUDA(cudaMemcpyAsync(h->map_n, mn, sizeof(mn), cudaMemcpyHostToDevice, h->stream));
    VLO_CUDA(cudaStreamSynchronize(h->stream));
    h->map_n_host[0] = n_corner; h->map_n_host[1] = n_surf;
    h->lm.mode = 1;                                   // static map: indices are the caller's, no cube window, no insertion
    for (int w = 0; w < 2; w++) {
        GridSource src = {};
        src.pts = h->map_pts[w]; src.pts_stride = 0; src.ring_off = nullptr; src.ring_off_stride = 0; src.ring_cnt = nullptr;
        src.ring_cnt_stride = 0; src.dense_start = nullptr; src.dense_start_stride = 0;
        src.n_dense = h->map_n; src.n_dense_stride = 0; src.n_dense_field = w == 0 ? 2 : 4; src.n_rings = 1; src.grid_scan = nullptr;
        int n = w == 0 ? n_corner : n_surf;
        int rc = vlo_grid_build(h, h->gs_map[w], src, 0, 1, std::max(n, 1)); if (rc) return rc;
    }
    return vlo_synchronize(h);
}

extern "C" int vlo_register_map(vlo_handle *h, const int *scans, int n, const float *seeds, vlo_result *out)
{
    if (!h || !scans || !seeds || !out || n < 1) return VLO_ERR_INVALID_ARG;
    if (h->cfg.max_map_points <= 0) { h->err = "handle created with max_map_points = 0"; return VLO_ERR_STATE; }
    if (n > h->cfg.max_scans) { h->err = "n exceeds max_scans"; return VLO_ERR_CAPACITY; }
    for (int k = 0; k < n; k++) if (scans[k] < 0 || scans[k] >= h->sb.n_scans) { h->err = "scan index outside the resident batch"; return VLO_ERR_INVALID_ARG; }
    cudaSetDevice(h->cfg.device);
    size_t need = sizeof(vlo_result) * (size_t)n + sizeof(int) * (size_t)n + sizeof(float) * 6 * (size_t)n;
    int rc = ensure_pinned2(h, need); if (rc) return rc;
    char *pp = (char *)h->pinned;
    vlo_result *pres = (vlo_result *)pp; pp += sizeof(vlo_result) * (size_t)n;
    int *ps = (int *)pp; pp += sizeof(int) * (size_t)n;
    float *pseed = (float *)pp;
    memcpy(ps, scans, sizeof(int) * (size_t)n); memcpy(pseed, seeds, sizeof(float) * 6 * (size_t)n);
    VLO_CUDA(cudaMemcpyAsync(h->map_scans, ps, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    VLO_CUDA(cudaMemcpyAsync(h->map_seed, pseed, sizeof(float) * 6 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    if (!h->lm.ds_valid) { rc = vlo_launch_stack_ds(h, 0, h->sb.n_scans); if (rc) return rc; h->lm.ds_valid = 1; }
    rc = vlo_launch_register_map(h, h->map_scans, n, h->map_seed); if (rc) return rc;
    VLO_CUDA(cudaMemcpyAsync(pres, h->map_result, sizeof(vlo_result) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    rc = vlo_synchronize(h); if (rc) return rc;
    memcpy(out, pres, sizeof(vlo_result) * (size_t)n);
    for (int k = 0; k < n; k++) vlo_finish_cov_host(&out[k], &h->cfg);
    h->last_n_map = n;
    int soft = VLO_OK;
    for (int k = 0; k < n; k++) if (out[k].status == VLO_SOFT_TOO_FEW_CORR) soft = VLO_SOFT_TOO_FEW_CORR;
    return soft;
}

extern "C" int vlo_register_map_enqueue(vlo_handle *h, const int *scans, int n, const float *seeds, vlo_result *out_pinned)
{
    if (!h || !scans || !seeds || !out_pinned || n < 1) return VLO_ERR_INVALID_ARG;
    if (h->cfg.max_map_points <= 0) { h->err = "handle created with max_map_points = 0"; return VLO_ERR_STATE; }
    if (n > h->cfg.max_scans) { h->err = "n exceeds max_scans"; return VLO_ERR_CAPACITY; }
    for (int k = 0; k < n; k++) if (scans[k] < 0 || scans[k] >= h->sb.n_scans) { h->err = "scan index outside the resident batch"; return VLO_ERR_INVALID_ARG; }
    cudaSetDevice(h->cfg.device);
    // the slot indices and seeds are small: a copy from pageable memory is staged by the runtime before the call returns, so
    // the caller's arrays may be reused at once and no staging of ours is in flight between calls
    VLO_CUDA(cudaMemcpyAsync(h->map_scans, scans, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    VLO_CUDA(cudaMemcpyAsync(h->map_seed, seeds, sizeof(float) * 6 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    int rc;
    if (!h->lm.ds_valid) { rc = vlo_launch_stack_ds(h, 0, h->sb.n_scans); if (rc) return rc; h->lm.ds_valid = 1; }
    rc = vlo_launch_register_map(h, h->map_scans, n, h->map_seed); if (rc) return rc;
    VLO_CUDA(cudaMemcpyAsync(out_pinned, h->map_result, sizeof(vlo_result) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    h->last_n_map = n;
    return VLO_OK;
}

extern "C" int vlo_results_finish(vlo_handle *h, vlo_result *out, int n)
{
    if (!h || !out || n < 0) return VLO_ERR_INVALID_ARG;
    int soft = VLO_OK;
    for (int k = 0; k < n; k++) { vlo_finish_cov_host(&out[k], &h->cfg); if (out[k].status == VLO_SOFT_TOO_FEW_CORR) soft = VLO_SOFT_TOO_FEW_CORR; }
    return soft;
}

extern "C" int vlo_map_get_correspondences(vlo_handle *h, int slot, int *corner_idx5, int *surf_idx5)
{
    // neighbour indices of the LAST executed association of slot `slot`
    if (!h || slot < 0 || slot >= h->last_n_map) return VLO_ERR_INVALID_ARG;
    int rc = vlo_synchronize(h); if (rc) return rc;
    int scan; VLO_CUDA(cudaMemcpy(&scan, h->map_scans + slot, sizeof(int), cudaMemcpyDeviceToHost));
    int cnt[8]; VLO_CUDA(cudaMemcpy(cnt, h->lm.ds_counts + scan * 8, sizeof(cnt), cudaMemcpyDeviceToHost));
    size_t qcap = (size_t)h->cap_lsharp + h->cfg.max_points;
    const int *base = h->map_idx5 + (size_t)slot * qcap * 5;
    if (corner_idx5 && cnt[2] > 0) VLO_CUDA(cudaMemcpy(corner_idx5, base, sizeof(int) * 5 * (size_t)cnt[2], cudaMemcpyDeviceToHost));
    if (surf_idx5 && cnt[4] > 0) VLO_CUDA(cudaMemcpy(surf_idx5, base + (size_t)cnt[2] * 5, sizeof(int) * 5 * (size_t)cnt[4], cudaMemcpyDeviceToHost));
    return VLO_OK;
}

extern "C" int vlo_map_knn(vlo_handle *h, int which, const float *q, int nq, int k, float max_d2, int *idx, float *d2)
{
    if (!h || which < 0 || which > 1 || !q || nq < 1 || !idx || !d2) return VLO_ERR_INVALID_ARG;
    if (h->cfg.max_map_points <= 0) { h->err = "handle created with max_map_points = 0"; return VLO_ERR_STATE; }
    cudaSetDevice(h->cfg.device);
    float4 *dq; int *di; float *dd;
    VLO_CUDA(cudaMalloc((void **)&dq, sizeof(float4) * (size_t)nq));
    VLO_CUDA(cudaMalloc((void **)&di, sizeof(int) * (size_t)nq * k));
    VLO_CUDA(cudaMalloc((void **)&dd, sizeof(float) * (size_t)nq * k));
    VLO_CUDA(cudaMemcpyAsync(dq, q, sizeof(float4) * (size_t)nq, cudaMemcpyHostToDevice, h->stream));
    int rc = vlo_grid_knn(h, h->gs_map[which], 0, dq, nq, k, max_d2, di, dd);
    if (rc == VLO_OK) {
        cudaMemcpyAsync(idx, di, sizeof(int) * (size_t)nq * k, cudaMemcpyDeviceToHost, h->stream);
        cudaMemcpyAsync(d2, dd, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, h->stream);
        rc = vlo_synchronize(h);
    }
    cudaFree(dq); cudaFree(di); cudaFree(dd);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// IMU
extern "C" int vlo_imu_preintegrate_batch(vlo_handle *h, const double *t, const double *acc, const double *gyro, int n_samples,
                                          const double *t0, const double *t1, const double *bias6, int n_factors, vlo_preint *out)
{
    if (!h || !t || !acc || !gyro || !t0 || !t1 || !out || n_samples < 0 || n_factors < 1) return VLO_ERR_INVALID_ARG;
    cudaSetDevice(h->cfg.device);
    size_t nd = (size_t)n_samples * 7 + (size_t)n_factors * 2 + 6;
    if (nd * sizeof(double) > h->imu_buf_bytes) {
        if (h->imu_buf) cudaFree(h->imu_buf);
        h->imu_buf = nullptr; h->imu_buf_bytes = 0;
        VLO_CUDA(cudaMalloc((void **)&h->imu_buf, nd * sizeof(double)));
        h->imu_buf_bytes = nd * sizeof(double);
    }
    if (n_factors > h->imu_out_cap) {
        if (h->imu_out) cudaFree(h->imu_out);
        h->imu_out = nullptr; h->imu_out_cap = 0;
        VLO_CUDA(cudaMalloc((void **)&h->imu_out, sizeof(vlo_preint) * (size_t)n_factors));
        h->imu_out_cap = n_factors;
    }
    double *d_t = h->imu_buf, *d_acc = d_t + n_samples, *d_gyro = d_acc + 3 * (size_t)n_samples, *d_t0 = d_gyro + 3 * (size_t)n_samples;
    double *d_t1 = d_t0 + n_factors, *d_bias = d_t1 + n_factors;
    double zero6[6] = { 0, 0, 0, 0, 0, 0 };
    if (n_samples) {
        VLO_CUDA(cudaMemcpyAsync(d_t, t, sizeof(double) * (size_t)n_samples, cudaMemcpyHostToDevice, h->stream));
        VLO_CUDA(cudaMemcpyAsync(d_acc, acc, sizeof(double) * 3 * (size_t)n_samples, cudaMemcpyHostToDevice, h->stream));
        VLO_CUDA(cudaMemcpyAsync(d_gyro, gyro, sizeof(double) * 3 * (size_t)n_samples, cudaMemcpyHostToDevice, h->stream));
    }
    VLO_CUDA(cudaMemcpyAsync(d_t0, t0, sizeof(double) * (size_t)n_factors, cudaMemcpyHostToDevice, h->stream));
    VLO_CUDA(cudaMemcpyAsync(d_t1, t1, sizeof(double) * (size_t)n_factors, cudaMemcpyHostToDevice, h->stream));
    VLO_CUDA(cudaMemcpyAsync(d_bias, bias6 ? bias6 : zero6, sizeof(double) * 6, cudaMemcpyHostToDevice, h->stream));
    int rc = vlo_launch_imu(h, d_t, d_acc, d_gyro, n_samples, d_t0, d_t1, d_bias, n_factors, h->imu_out); if (rc) return rc;
    VLO_CUDA(cudaMemcpyAsync(out, h->imu_out, sizeof(vlo_preint) * (size_t)n_factors, cudaMemcpyDeviceToHost, h->stream));
    return vlo_synchronize(h);
}

// ---------------------------------------------------------------------------------------------
// online tick: ping-pong between two resident scan slots
static void euler_to_M(const float *T, double M[4][4])
{
    double sx = sin(T[0]), cx = cos(T[0]), sy = sin(T[1]), cy = cos(T[1]), sz = sin(T[2]), cz = cos(T[2]);
    M[0][0] = cy * cz + sy * sx * sz; M[0][1] = -cy * sz + sy * sx * cz; M[0][2] = sy * cx;
    M[1][0] = cx * sz;                M[1][1] = cx * cz;                 M[1][2] = -sx;
    M[2][0] = -sy * cz + cy * sx * sz; M[2][1] = sy * sz + cy * sx * cz; M[2][2] = cy * cx;
    M[0][3] = T[3]; M[1][3] = T[4]; M[2][3] = T[5];
    M[3][0] = M[3][1] = M[3][2] = 0; M[3][3] = 1;
}
static void M_to_euler(const double M[4][4], float *T)
{
    T[0] = (float)(-asin(std::max(-1.0, std::min(1.0, M[1][2]))));
    T[1] = (float)atan2(M[0][2], M[2][2]);
    T[2] = (float)atan2(M[1][0], M[1][1]);
    T[3] = (float)M[0][3]; T[4] = (float)M[1][3]; T[5] = (float)M[2][3];
}
static void M_mul(const double A[4][4], const double B[4][4], double C[4][4])
{
    double t[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += A[i][k] * B[k][j]; t[i][j] = s; }
    memcpy(C, t, sizeof(t));
}
static void M_inv(const double A[4][4], double B[4][4])
{
    double t[4][4] = { { 0 } };
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[i][j] = A[j][i];
    for (int i = 0; i < 3; i++) t[i][3] = -(t[i][0] * A[0][3] + t[i][1] * A[1][3] + t[i][2] * A[2][3]);
    t[3][3] = 1;
    memcpy(B, t, sizeof(t));
}

extern "C" int vlo_online_reset(vlo_handle *h)
{
    if (!h) return VLO_ERR_INVALID_ARG;
    h->online_have_last = 0; h->online_slot = 0; h->online_ticks = 0;
    memset(h->online_T, 0, sizeof(h->online_T)); memset(h->online_sum, 0, sizeof(h->online_sum));
    memset(h->online_map_bef, 0, sizeof(h->online_map_bef)); memset(h->online_map_aft, 0, sizeof(h->online_map_aft));
    return VLO_OK;
}

extern "C" int vlo_online_pose(vlo_handle *h, float *sum6, float *mapped6)
{
    if (!h) return VLO_ERR_INVALID_ARG;
    if (sum6) memcpy(sum6, h->online_sum, sizeof(float) * 6);
    if (mapped6) memcpy(mapped6, h->online_map_aft, sizeof(float) * 6);
    return VLO_OK;
}

extern "C" int vlo_online_set_map_pose(vlo_handle *h, const float *pose6)
{
    if (!h || !pose6) return VLO_ERR_INVALID_ARG;
    memcpy(h->online_map_aft, pose6, sizeof(float) * 6);
    memcpy(h->online_sum, pose6, sizeof(float) * 6);
    memcpy(h->online_map_bef, pose6, sizeof(float) * 6);
    return VLO_OK;
}

static int process_scan_impl(vlo_handle *h, const float *raw, int n_points, int stride, const int *xyz_off, vlo_result *odom, vlo_result *mapped);

extern "C" int vlo_process_scan(vlo_handle *h, const float *raw, int n_points, int stride, double stamp, vlo_result *odom, vlo_result *mapped)
{
    (void)stamp;
    const int off[3] = { 0, 1, 2 };
    return process_scan_impl(h, raw, n_points, stride, off, odom, mapped);
}

extern "C" int vlo_process_scan_pc2(vlo_handle *h, const void *data, int n_points, int point_step, int x_offset, int y_offset, int z_offset,
                                    double stamp, vlo_result *odom, vlo_result *mapped)
{
    (void)stamp;
    if (!h || point_step < 12 || (point_step & 3) || (x_offset & 3) || (y_offset & 3) || (z_offset & 3) || x_offset < 0 || y_offset < 0 ||
        z_offset < 0 || x_offset + 4 > point_step || y_offset + 4 > point_step || z_offset + 4 > point_step ||
        x_offset == y_offset || x_offset == z_offset || y_offset == z_offset) {
        if (h) h->err = "PointCloud2 layout: point_step and the x/y/z offsets must be multiples of 4 inside the point (FLOAT32 fields)";
        return VLO_ERR_INVALID_ARG;
    }
    const int off[3] = { x_offset / 4, y_offset / 4, z_offset / 4 };
    return process_scan_impl(h, (const float *)data, n_points, point_step / 4, off, odom, mapped);
}

static int process_scan_impl(vlo_handle *h, const float *raw, int n_points, int stride, const int *xyz_off, vlo_result *odom, vlo_result *mapped)
{
    if (!h || !raw || n_points < 0 || stride < 3) return VLO_ERR_INVALID_ARG;
    if (h->cfg.max_scans < 2) { h->err = "online mode needs max_scans >= 2"; return VLO_ERR_STATE; }
    if (n_points > h->cfg.max_points || (size_t)n_points * stride > (size_t)h->cfg.max_points * 4) { h->err = "scan exceeds max_points"; return VLO_ERR_CAPACITY; }
    cudaSetDevice(h->cfg.device);
    ScanBatchDev &sb = h->sb;
    const int cur = h->online_slot, last = 1 - cur;
    const size_t N = (size_t)h->cfg.max_points;
    int rc = ensure_pinned2(h, std::max<size_t>(sizeof(vlo_result) * 2 + 256, (size_t)1 << 12)); if (rc) return rc;
    int *poff = (int *)((char *)h->pinned + sizeof(vlo_result) * 2);
    poff[0] = (int)((cur * N * 4 + stride - 1) / stride); poff[1] = n_points;
    // slot `cur` owns raw_owned[cur*N*4 .. ); its begin index is expressed in points of this stride
    float *dst = sb.raw_owned + (size_t)poff[0] * stride;
    VLO_CUDA(cudaMemcpyAsync(dst, raw, sizeof(float) * (size_t)n_points * stride, cudaMemcpyHostToDevice, h->stream));
    VLO_CUDA(cudaMemcpyAsync(sb.raw_offset + 2 * cur, poff, sizeof(int) * 2, cudaMemcpyHostToDevice, h->stream));
    sb.raw = sb.raw_owned; sb.stride = stride; sb.n_scans = 2; sb.scan_first = cur; sb.scan_count = 1;
    sb.xyz_off[0] = xyz_off[0]; sb.xyz_off[1] = xyz_off[1]; sb.xyz_off[2] = xyz_off[2];
    rc = vlo_launch_organise(h); if (rc) return rc;
    rc = vlo_launch_extract(h); if (rc) return rc;
    h->map_qmax = 0;
    vlo_result *pres = (vlo_result *)h->pinned;
    int soft = VLO_OK;
    // a capacity report (oversized ring, map full) concerns this sweep only: the tick's state (slots, transformSum, map pose)
    // still advances, the code is returned at the end -- the next tick then registers against THIS sweep, as the
    // reference's nodelets would
    int deferred = VLO_OK;
    bool did_odom = false;
    if (h->online_have_last) {
        int *pl = poff + 2;                                      // pinned staging (after the offsets)
        float *seedT = (float *)(pl + 2);
        pl[0] = last; pl[1] = cur;
        memcpy(seedT, h->online_T, sizeof(float) * 6);          // seed = previous transform (constant velocity)
        memcpy(seedT + 6, h->online_T, sizeof(float) * 6);      // last sweep is moved to its end with the same transform
        VLO_CUDA(cudaMemcpyAsync(h->pair_last, &pl[0], sizeof(int), cudaMemcpyHostToDevice, h->stream));
        VLO_CUDA(cudaMemcpyAsync(h->pair_cur, &pl[1], sizeof(int), cudaMemcpyHostToDevice, h->stream));
        VLO_CUDA(cudaMemcpyAsync(h->pair_seed, seedT, sizeof(float) * 6, cudaMemcpyHostToDevice, h->stream));
        VLO_CUDA(cudaMemcpyAsync(h->pair_last_T, seedT + 6, sizeof(float) * 6, cudaMemcpyHostToDevice, h->stream));
        // `last` was already moved to its sweep end right after its own registration (below); the first sweep has no
        // estimate of its own motion and stays untouched, as upstream stores it
        rc = vlo_launch_register_pairs(h, 1, h->pair_seed, nullptr, last); if (rc) return rc;
        VLO_CUDA(cudaMemcpyAsync(pres, h->pair_result, sizeof(vlo_result), cudaMemcpyDeviceToHost, h->stream));
        // transformToEnd of this sweep's target clouds with its own transform (device-resident, no host round trip):
        // they are the next tick's `last` clouds and the stack LaserMapping receives
        if (h->cfg.deskew) { rc = vlo_launch_to_end(h, h->pair_cur, h->pair_T, 1); if (rc) return rc; }
        rc = vlo_synchronize(h);
        if (rc == VLO_ERR_CAPACITY) deferred = rc; else if (rc) return rc;
        h->last_n_pairs = 1;
        vlo_finish_cov_host(pres, &h->cfg);
        if (pres->status == VLO_OK) memcpy(h->online_T, pres->transform, sizeof(float) * 6);
        else soft = pres->status;
        vlo_accumulate_pose(h->online_sum, h->online_T, 1.0f, h->online_sum);
        if (odom) *odom = *pres;
        did_odom = true;
    } else {
        rc = vlo_synchronize(h);
        if (rc == VLO_ERR_CAPACITY) deferred = rc; else if (rc) return rc;
    }
    if (!did_odom && odom) { memset(odom, 0, sizeof(*odom)); for (int a = 0; a < 6; a++) odom->P[a * 7] = 1.0f; odom->status = VLO_SOFT_TOO_FEW_CORR; }
    if (mapped) {
        memset(mapped, 0, sizeof(*mapped)); mapped->status = VLO_SOFT_TOO_FEW_CORR;
        for (int a = 0; a < 6; a++) mapped->P[a * 7] = 1.0f;
        // upstream: laserOdometry hands a sweep to laserMapping when ioRatio < 2 or frameCount % ioRatio == 1
        const bool due = h->cfg.io_ratio < 2 || (h->online_ticks % h->cfg.io_ratio) == 1;
        if (h->cfg.max_map_points > 0 && due) {
            // transformAssociateToMap: seed = aft * bef^-1 * sum
            double Ma[4][4], Mb[4][4], Ms[4][4], Mbi[4][4], Mt[4][4];
            euler_to_M(h->online_map_aft, Ma); euler_to_M(h->online_map_bef, Mb); euler_to_M(h->online_sum, Ms);
            M_inv(Mb, Mbi); M_mul(Ma, Mbi, Mt); M_mul(Mt, Ms, Mt);
            float seed[6]; M_to_euler(Mt, seed);
            rc = vlo_launch_stack_ds(h, cur, 1); if (rc) return rc;
            h->lm.ds_valid = 1;
            int mrc = VLO_OK;
            if (h->lm.mode == 1) {
                // static map (vlo_map_build): registration only
                if (h->map_n_host[0] > 10 && h->map_n_host[1] > 100) { int sc = cur; mrc = vlo_register_map(h, &sc, 1, seed, mapped); }
            } else {
                // maintained map: the whole BasicLaserMapping::process (sub-map, optimisation, insertion)
                mrc = vlo_map_process(h, cur, seed, mapped, nullptr);
            }
            if (mrc == VLO_ERR_CAPACITY) deferred = mrc;        // the record in `mapped` is valid (vlo_map_process fills it)
            else if (mrc < 0) return mrc;
            if (mapped->status == VLO_OK || h->lm.mode != 1) {
                // transformUpdate: upstream stores the pose whether or not the optimisation ran
                memcpy(h->online_map_aft, mapped->transform, sizeof(float) * 6); memcpy(h->online_map_bef, h->online_sum, sizeof(float) * 6);
            }
            if (mapped->status != VLO_OK) soft = mapped->status;
        }
    }
    h->online_have_last = 1; h->online_slot = last; h->online_ticks++;
    return deferred ? deferred : soft;
}

// ---------------------------------------------------------------------------------------------
// host helpers (float64 / float32 scalar code; not on the device path)
static void quat_mul(const double *a, const double *b, double *o)
{
    double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}

extern "C" void vlo_pose_diff(const double *b7, const double *a7, double *out7)
{
    // SensorManagerRos.cpp:142-151: dxr = q1^-1 * (x2 - x1);  qr = q2 * q1^-1
    const double *q1 = b7 + 3, *q2 = a7 + 3;
    double n2 = q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3];
    double qi[4] = { q1[0] / n2, -q1[1] / n2, -q1[2] / n2, -q1[3] / n2 };
    double v[3] = { a7[0] - b7[0], a7[1] - b7[1], a7[2] - b7[2] };
    double u[3] = { qi[1], qi[2], qi[3] };
    double uv[3] = { u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0] };
    for (int i = 0; i < 3; i++) uv[i] += uv[i];
    double uuv[3] = { u[1] * uv[2] - u[2] * uv[1], u[2] * uv[0] - u[0] * uv[2], u[0] * uv[1] - u[1] * uv[0] };
    for (int i = 0; i < 3; i++) out7[i] = v[i] + qi[0] * uv[i] + uuv[i];
    quat_mul(q2, qi, out7 + 3);
}

static float det3h(const float *H, int o)
{
#define M_(r, c) H[(o + (r)) * 6 + (o + (c))]
    float a = M_(0, 0) * (M_(1, 1) * M_(2, 2) - M_(1, 2) * M_(2, 1));
    float b = M_(0, 1) * (M_(1, 0) * M_(2, 2) - M_(1, 2) * M_(2, 0));
    float c = M_(0, 2) * (M_(1, 0) * M_(2, 1) - M_(1, 1) * M_(2, 0));
#undef M_
    return (a - b) + c;
}

extern "C" int vlo_dopt_gate(const float *H, double rot_thr, double trans_thr, float *logdet_rot, float *logdet_trans)
{
    // degerate_odometry_filter.cpp:32-46: rotation = block(3,3), translation = block(0,0); NaN passes
    float rot = logf(det3h(H, 3)), trans = logf(det3h(H, 0));
    if (logdet_rot) *logdet_rot = rot;
    if (logdet_trans) *logdet_trans = trans;
    return ((double)rot < rot_thr || (double)trans < trans_thr) ? 0 : 1;
}

static void euler_to_R(double rx, double ry, double rz, double R[3][3])
{
    double sx = sin(rx), cx = cos(rx), sy = sin(ry), cy = cos(ry), sz = sin(rz), cz = cos(rz);
    R[0][0] = cy * cz + sy * sx * sz; R[0][1] = -cy * sz + sy * sx * cz; R[0][2] = sy * cx;
    R[1][0] = cx * sz;                R[1][1] = cx * cz;                 R[1][2] = -sx;
    R[2][0] = -sy * cz + cy * sx * sz; R[2][1] = sy * sz + cy * sx * cz; R[2][2] = cy * cx;
}

extern "C" void vlo_accumulate_pose(const float *sum_in, const float *T, float fudge, float *sum_out)
{
    // transformSum <- transformSum (+) T with R = Ry Rx Rz:  R' = R_sum R_T^-1,  p' = p_sum - R' t
    double Rs[3][3], Rt[3][3], Rn[3][3];
    euler_to_R(sum_in[0], sum_in[1], sum_in[2], Rs);
    euler_to_R(T[0], (double)T[1] * fudge, T[2], Rt);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += Rs[i][k] * Rt[j][k];
        Rn[i][j] = s;
    }
    double t[3] = { T[3], T[4], (double)T[5] * fudge };
    sum_out[0] = (float)(-asin(Rn[1][2]));
    sum_out[1] = (float)atan2(Rn[0][2], Rn[2][2]);
    sum_out[2] = (float)atan2(Rn[1][0], Rn[1][1]);
    for (int i = 0; i < 3; i++) {
        double s = 0; for (int k = 0; k < 3; k++) s += Rn[i][k] * t[k];
        sum_out[3 + i] = (float)((double)sum_in[3 + i] - s);
    }
}
