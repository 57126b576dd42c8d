// C-ABI of libvlo.so (include/vlo.h): handle lifecycle, uploads, stage launchers, copy-backs.
// Host logic only; every computation on the hot path is a kernel in k*.cu.  No CPU fallback: with
// no usable CUDA device vlo_create fails with VLO_ERR_NO_DEVICE.
#include "vlo_internal.cuh"
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <algorithm>

static const char *kVersion = "vlo-b200 0.1 (sm_100a)";

extern "C" const char *vlo_version(void) { return kVersion; }

extern "C" void vlo_default_config(vlo_config *c)
{
    memset(c, 0, sizeof(*c));
    c->max_scans = 2; c->max_points = 32768; c->max_ring_points = 2048; c->max_map_points = 0;
    c->max_imu_factors = 0; c->max_imu_samples = 0; c->device = 0;
    c->scan_period = 0.1f; c->n_rings = 16; c->lower_deg = -15.0f; c->upper_deg = 15.0f;
    c->feature_regions = 6; c->curvature_region = 5; c->max_corner_sharp = 2; c->max_corner_less_sharp = 20;
    c->max_surface_flat = 4; c->surface_curvature_threshold = 0.1f; c->less_flat_filter_size = 0.2f;
    c->odom_max_iterations = 25; c->odom_delta_t_abort = 0.05f; c->odom_delta_r_abort = 0.05f; c->odom_degen_eig = 30.0f;
    c->deskew = 1; c->odom_forward_bound_quirk = 0;
    c->map_max_iterations = 10; c->map_delta_t_abort = 0.05f; c->map_delta_r_abort = 0.05f; c->map_degen_eig = 40.0f;
    c->map_cell_size = 1.0625f; c->odom_cell_size = 1.0f; c->odom_corner_cell_size = 5.0f;
    c->dopt_rot_threshold = 11.5f; c->dopt_trans_threshold = 28.9f;
    c->cov_accel = 1e-6; c->cov_gyro = 1e-6; c->cov_integration = 1e-8; c->cov_bias_acc = 1e-4;
    c->cov_bias_omega = 1e-6; c->cov_bias_acc_omega_int = 1e-4;
    c->corner_filter_size = 0.2f; c->surface_filter_size = 0.4f; c->map_cube_size = 10.0f;
    c->map_dims[0] = 101; c->map_dims[1] = 51; c->map_dims[2] = 101;
    c->map_start_cubes[0] = 50; c->map_start_cubes[1] = 25; c->map_start_cubes[2] = 50;
    c->n_neighbor_cubes = 5; c->io_ratio = 2; c->hessian_order = 0;
    c->undistort_input_cloud = 0;
    c->rotate_input = 0; c->input_rotation[0] = c->input_rotation[1] = c->input_rotation[2] = 0.0f; c->ring_field = -1; c->ring_field_type = 0;
}

extern "C" int vlo_set_lidar(vlo_config *c, const char *name)
{
    if (!strcmp(name, "VLP-16")) { c->n_rings = 16; c->lower_deg = -15.0f; c->upper_deg = 15.0f; return VLO_OK; }
    if (!strcmp(name, "HDL-32")) { c->n_rings = 32; c->lower_deg = -30.67f; c->upper_deg = 10.67f; return VLO_OK; }
    if (!strcmp(name, "HDL-64E")) { c->n_rings = 64; c->lower_deg = -24.9f; c->upper_deg = 2.0f; return VLO_OK; }
    // the fork's additional presets named in loam_params.yaml:22 (values: the sensors' data-sheet fields of view; the fork's
    // own numbers are not in /root/reference)
    if (!strcmp(name, "O1-16")) { c->n_rings = 16; c->lower_deg = -16.611f; c->upper_deg = 16.611f; return VLO_OK; }
    if (!strcmp(name, "O1-64")) { c->n_rings = 64; c->lower_deg = -16.611f; c->upper_deg = 16.611f; return VLO_OK; }
    if (!strcmp(name, "Bperl-32")) { c->n_rings = 32; c->lower_deg = 2.3125f; c->upper_deg = 89.5f; return VLO_OK; }
    return VLO_ERR_INVALID_ARG;
}

template <typename T> static cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)); }

#define HALLOC(ptr, n) do { cudaError_t e_ = dalloc(&(ptr), (n)); if (e_ != cudaSuccess) { \
    h->err = std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_); vlo_destroy(h); return VLO_ERR_CUDA; } } while (0)

static int ensure_pinned(vlo_handle *h, size_t bytes)
{
    if (bytes <= h->pinned_bytes) return VLO_OK;
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr; h->pinned_bytes = 0;
    VLO_CUDA(cudaMallocHost(&h->pinned, bytes));
    h->pinned_bytes = bytes;
    return VLO_OK;
}

static int alloc_gridset(vlo_handle *h, GridSet &g, int n_grids, int max_pts, float cell)
{
    int ts = 1024; while (ts <= max_pts) ts <<= 1;      // at least one empty slot: every probe sequence terminates
    g.cell = cell; g.inv_cell = 1.0f / cell; g.ts = ts; g.max_pts = max_pts; g.G = n_grids;
    cudaError_t e;
    if ((e = dalloc(&g.keys, (size_t)n_grids * ts)) != cudaSuccess || (e = dalloc(&g.cnt, (size_t)n_grids * ts)) != cudaSuccess ||
        (e = dalloc(&g.start, (size_t)n_grids * (ts + 1))) != cudaSuccess || (e = dalloc(&g.sorted, (size_t)n_grids * max_pts)) != cudaSuccess ||
        (e = dalloc(&g.bsum, (size_t)n_grids * ((ts + GRID_SCAN_BLOCK - 1) / GRID_SCAN_BLOCK))) != cudaSuccess) {
        h->err = std::string("cudaMalloc grid: ") + cudaGetErrorString(e); return VLO_ERR_CUDA;
    }
    return VLO_OK;
}

static void free_gridset(GridSet &g)
{
    cudaFree(g.keys); cudaFree(g.cnt); cudaFree(g.start); cudaFree(g.sorted); cudaFree(g.bsum);
    memset(&g, 0, sizeof(g));
}

extern "C" int vlo_create(const vlo_config *cfg, vlo_handle **out)
{
    if (!cfg || !out) return VLO_ERR_INVALID_ARG;
    *out = nullptr;
    const vlo_config &c = *cfg;
    if (c.max_scans < 1 || c.max_points < 1 || c.n_rings < 1 || c.n_rings > VLO_MAX_RINGS ||
        c.feature_regions < 1 || c.feature_regions > VLO_MAX_REGIONS || c.curvature_region < 1 || c.curvature_region > 15 ||
        c.max_ring_points < 32 || c.max_ring_points > 4096 || c.max_corner_sharp < 0 || c.max_corner_less_sharp < c.max_corner_sharp ||
        c.max_corner_less_sharp > 255 || c.max_surface_flat < 0 || c.max_surface_flat > 255 || !(c.upper_deg > c.lower_deg) ||
        !(c.less_flat_filter_size > 0.f) || !(c.scan_period > 0.f))
        return VLO_ERR_INVALID_ARG;
    if (c.undistort_input_cloud) return VLO_ERR_UNSUPPORTED;      // see vlo.h: not implemented, never silently mapped onto something else
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || c.device >= ndev) return VLO_ERR_NO_DEVICE;
    if (cudaSetDevice(c.device) != cudaSuccess) return VLO_ERR_NO_DEVICE;
    vlo_handle *h = new vlo_handle();
    h->cfg = c; h->launches = 0; h->pinned = nullptr; h->pinned_bytes = 0; h->upload_pinned = nullptr;
    memset(&h->sb, 0, sizeof(h->sb)); memset(&h->lm, 0, sizeof(h->lm));
    memset(&h->gs_corner, 0, sizeof(GridSet)); memset(&h->gs_surf, 0, sizeof(GridSet)); memset(&h->segs, 0, sizeof(SegSet)); memset(h->gs_map, 0, sizeof(h->gs_map));
    h->map_pts[0] = h->map_pts[1] = nullptr; h->map_n = nullptr; h->map_n_host[0] = h->map_n_host[1] = 0;
    h->grids_valid = 0; h->scan_index_grid = 0; h->trace = 0; h->pair_last_T = nullptr;
    h->status_word = nullptr; h->pair_T = h->pair_seed = nullptr; h->pair_last = h->pair_cur = h->pair_state = nullptr;
    h->pair_cidx = h->pair_sidx = h->pair_trace = nullptr; h->pair_result = nullptr;
    h->map_partials = nullptr; h->map_idx5 = nullptr; h->map_T = h->map_seed = nullptr; h->map_state = h->map_ncorr = h->map_scans = h->map_done = nullptr;
    h->k1_smem_configured = h->k1c_smem_configured = 0; h->dev_sms = h->k5_occ_assoc = h->k5_occ_lin = h->k3_gn_configured = h->k3a_ctas = 0; h->bag_ctx = nullptr;
    { const char *e0 = getenv("VLO_K0_SUB"), *e1 = getenv("VLO_K1_SUB"); h->k0_sub = e0 ? atoi(e0) : 0; h->k1_sub = e1 ? atoi(e1) : 0; }
    h->map_result = nullptr; h->coop_resident = 0; h->map_qmax = 0; h->last_n_map = 0; h->last_n_pairs = 0;
    h->imu_buf = nullptr; h->imu_buf_bytes = 0; h->imu_out = nullptr; h->imu_out_cap = 0;
    h->online_have_last = 0; h->online_slot = 0; h->prof_enabled = 0; h->prof_used = 0;
    memset(h->online_T, 0, sizeof(h->online_T)); memset(h->online_sum, 0, sizeof(h->online_sum));
    memset(h->online_map_bef, 0, sizeof(h->online_map_bef)); memset(h->online_map_aft, 0, sizeof(h->online_map_aft)); h->online_ticks = 0;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return VLO_ERR_CUDA; }
    const int B = c.max_scans, N = c.max_points, R = c.n_rings, NR = c.feature_regions;
    h->tiles_per_scan = (N + K0_TILE - 1) / K0_TILE;
    h->cap_sharp = R * NR * std::max(c.max_corner_sharp, 1);
    h->cap_lsharp = R * NR * std::max(c.max_corner_less_sharp, 1);
    h->cap_flat = R * NR * std::max(c.max_surface_flat, 1);
    ScanBatchDev &sb = h->sb;
    HALLOC(h->status_word, 1);
    cudaMemset(h->status_word, 0, sizeof(int));
    HALLOC(sb.raw_owned, (size_t)B * N * 4 + 64);
    HALLOC(sb.raw_offset, (size_t)B * 2);
    HALLOC(sb.first_half, (size_t)B); HALLOC(sb.ori_bounds, (size_t)B * 2);
    HALLOC(sb.tile_hist, (size_t)B * R * h->tiles_per_scan);
    HALLOC(sb.ring_of, (size_t)B * N); HALLOC(sb.ori_of, (size_t)B * N);
    HALLOC(sb.cloud, (size_t)B * N); HALLOC(sb.ring_start, (size_t)B * (VLO_MAX_RINGS + 1)); HALLOC(sb.src_index, (size_t)B * N);
    HALLOC(sb.label, (size_t)B * N); HALLOC(sb.curvature, (size_t)B * N); HALLOC(sb.picked, (size_t)B * N);
    HALLOC(sb.slot_sharp, (size_t)B * h->cap_sharp); HALLOC(sb.slot_lsharp, (size_t)B * h->cap_lsharp);
    HALLOC(sb.slot_flat, (size_t)B * h->cap_flat); HALLOC(sb.slot_cnt, (size_t)B * R * NR * 4);
    HALLOC(sb.lflat_slotted, (size_t)B * N); HALLOC(sb.lflat_cnt, (size_t)B * R);
    HALLOC(sb.counts, (size_t)B * 8);
    HALLOC(sb.sharp_idx, (size_t)B * h->cap_sharp); HALLOC(sb.lsharp_idx, (size_t)B * h->cap_lsharp); HALLOC(sb.flat_idx, (size_t)B * h->cap_flat);
    HALLOC(sb.sharp_pts, (size_t)B * h->cap_sharp); HALLOC(sb.lsharp_pts, (size_t)B * h->cap_lsharp); HALLOC(sb.flat_pts, (size_t)B * h->cap_flat);
    HALLOC(sb.lflat_pts, (size_t)B * N);
    HALLOC(sb.lsharp_ring_start, (size_t)B * (VLO_MAX_RINGS + 1)); HALLOC(sb.lflat_ring_start, (size_t)B * (VLO_MAX_RINGS + 1));
    cudaMemset(sb.counts, 0, (size_t)B * 8 * sizeof(int));
    // registration workspace: at most one pair per resident scan
    h->max_pairs = B;
    HALLOC(h->pair_T, (size_t)B * 6); HALLOC(h->pair_seed, (size_t)B * 6);
    HALLOC(h->pair_last, (size_t)B); HALLOC(h->pair_cur, (size_t)B); HALLOC(h->pair_state, (size_t)B * 4 + 64);       // + one work counter per association round
    HALLOC(h->pair_cidx, (size_t)B * h->cap_sharp * 2); HALLOC(h->pair_sidx, (size_t)B * h->cap_flat * 3);
    HALLOC(h->pair_trace, (size_t)B * 5 * (h->cap_sharp * 2 + h->cap_flat * 3));
    HALLOC(h->pair_result, (size_t)B);
    HALLOC(h->pair_last_T, (size_t)B * 6);
    { int rc = alloc_gridset(h, h->gs_corner, B, h->cap_lsharp, c.odom_corner_cell_size > 0.f ? c.odom_corner_cell_size : 5.0f); if (rc) { vlo_destroy(h); return rc; } }
    { int rc = alloc_gridset(h, h->gs_surf, B, N, c.odom_cell_size); if (rc) { vlo_destroy(h); return rc; } }
    // ring-segment box indices of the scan-to-scan target clouds (segbox.cuh): SEG_PTS-point arcs, at most one partial arc per ring
    for (int w = 0; w < 2; w++) {
        SegSet &ss = h->segs;
        ss.max_seg[w] = (w == 0 ? h->cap_lsharp : N) / SEG_PTS + R + 1;
        ss.max_coarse[w] = ss.max_seg[w] / 32 + 1;
        HALLOC(ss.fbox[w], (size_t)B * ss.max_seg[w] * 2); HALLOC(ss.mbox[w], (size_t)B * ss.max_seg[w] * 2); HALLOC(ss.cbox[w], (size_t)B * ss.max_coarse[w] * 2);
        HALLOC(ss.perm[w], (size_t)B * ss.max_seg[w]); HALLOC(ss.seg_ring[w], (size_t)B * (VLO_MAX_RINGS + 1)); HALLOC(ss.nseg[w], (size_t)B);
        HALLOC(ss.prange[w], (size_t)B * (w == 0 ? h->cap_lsharp : N));
    }
    HALLOC(h->map_n, 8);
    cudaMemset(h->map_n, 0, 8 * sizeof(int));
    if (c.max_map_points > 0) {
        for (int w = 0; w < 2; w++) {
            HALLOC(h->map_pts[w], (size_t)c.max_map_points);
            int rc = alloc_gridset(h, h->gs_map[w], 1, c.max_map_points, c.map_cell_size); if (rc) { vlo_destroy(h); return rc; }
        }
        int qcap = h->cap_lsharp + N;
        HALLOC(h->map_partials, (size_t)B * ((qcap + 31) / 32 + 8) * VLO_NTERM);
        HALLOC(h->map_idx5, (size_t)B * qcap * 5);
        HALLOC(h->map_T, (size_t)B * 6); HALLOC(h->map_seed, (size_t)B * 6); HALLOC(h->map_state, (size_t)B * 4);
        HALLOC(h->map_ncorr, (size_t)B * 2); HALLOC(h->map_done, (size_t)B * 2 + 2 * (size_t)std::max(1, c.map_max_iterations) + 1); HALLOC(h->map_scans, (size_t)B); HALLOC(h->map_result, (size_t)B);
        { int rc = vlo_lm_alloc(h); if (rc) { vlo_destroy(h); return rc; } }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { h->err = "device sync after allocation failed"; vlo_destroy(h); return VLO_ERR_CUDA; }
    *out = h;
    return VLO_OK;
}

extern "C" void vlo_destroy(vlo_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    ScanBatchDev &sb = h->sb;
    void *ptrs[] = { sb.raw_owned, sb.raw_offset, sb.first_half, sb.ori_bounds, sb.tile_hist, sb.ring_of, sb.ori_of, sb.cloud, sb.ring_start, sb.src_index,
                     sb.label, sb.curvature, sb.picked, sb.slot_sharp, sb.slot_lsharp, sb.slot_flat, sb.slot_cnt, sb.lflat_slotted,
                     sb.lflat_cnt, sb.counts, sb.sharp_idx, sb.lsharp_idx, sb.flat_idx, sb.sharp_pts, sb.lsharp_pts, sb.flat_pts,
                     sb.lsharp_ring_start, sb.lflat_ring_start, sb.lflat_pts, h->status_word, h->pair_T, h->pair_seed, h->pair_last,
                     h->pair_cur, h->pair_state, h->pair_cidx, h->pair_sidx, h->pair_trace, h->pair_result, h->pair_last_T, h->map_n,
                     h->map_pts[0], h->map_pts[1], h->map_partials, h->map_idx5, h->map_T, h->map_seed, h->map_state, h->map_ncorr, h->map_done,
                     h->map_scans, h->map_result, h->imu_buf, h->imu_out };
    for (void *p : ptrs) if (p) cudaFree(p);
    vlo_lm_free(h);
    vlo_bag_free(h);
    for (int w = 0; w < 2; w++) { cudaFree(h->segs.fbox[w]); cudaFree(h->segs.mbox[w]); cudaFree(h->segs.cbox[w]); cudaFree(h->segs.perm[w]); cudaFree(h->segs.seg_ring[w]); cudaFree(h->segs.nseg[w]); cudaFree(h->segs.prange[w]); }
    free_gridset(h->gs_corner); free_gridset(h->gs_surf); free_gridset(h->gs_map[0]); free_gridset(h->gs_map[1]);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->upload_pinned) { cudaFreeHost(h->upload_pinned); cudaEventDestroy(h->upload_ev[0]); cudaEventDestroy(h->upload_ev[1]); }
    for (auto &e : h->prof_events) cudaEventDestroy(e);
    cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char *vlo_last_error(const vlo_handle *h) { return h ? h->err.c_str() : "null handle"; }
extern "C" long long vlo_launch_count(const vlo_handle *h) { return h ? h->launches : 0; }

extern "C" int vlo_synchronize(vlo_handle *h)
{
    if (!h) return VLO_ERR_INVALID_ARG;
    VLO_CUDA(cudaStreamSynchronize(h->stream));
    int st = 0;
    VLO_CUDA(cudaMemcpy(&st, h->status_word, sizeof(int), cudaMemcpyDeviceToHost));
    if (!st) return VLO_OK;
    // both conditions are reported ONCE and cleared: the oversized ring was skipped for that sweep only (the reference
    // has no per-ring cap and carries on with the next sweep), the map keeps working with the voxels it has
    const int zero = 0;
    VLO_CUDA(cudaMemcpy(h->status_word, &zero, sizeof(int), cudaMemcpyHostToDevice));
    if (st & 1) { h->err = "a ring holds more points than max_ring_points (that ring yielded no features)"; return VLO_ERR_CAPACITY; }
    h->err = "the maintained map is full (max_map_points): new voxels were dropped";
    return VLO_ERR_CAPACITY;
}

extern "C" int vlo_scans_upload_pc2(vlo_handle *h, const void *data, const int *offsets, int n_scans, int point_step,
                                    int x_offset, int y_offset, int z_offset, int on_device)
{
    // sensor_msgs/PointCloud2: little-endian float32 fields at byte offsets inside points of point_step bytes
    if (!h || point_step < 12 || (point_step & 3) || (x_offset & 3) || (y_offset & 3) || (z_offset & 3) || x_offset < 0 || y_offset < 0 ||
        z_offset < 0 || x_offset + 4 > point_step || y_offset + 4 > point_step || z_offset + 4 > point_step ||
        x_offset == y_offset || x_offset == z_offset || y_offset == z_offset) {
        if (h) h->err = "PointCloud2 layout: point_step and the x/y/z offsets must be multiples of 4 inside the point (FLOAT32 fields)";
        return VLO_ERR_INVALID_ARG;
    }
    int rc = vlo_scans_upload(h, (const float *)data, offsets, n_scans, point_step / 4, on_device);
    if (rc == VLO_OK) { h->sb.xyz_off[0] = x_offset / 4; h->sb.xyz_off[1] = y_offset / 4; h->sb.xyz_off[2] = z_offset / 4; }
    return rc;
}

extern "C" int vlo_scans_upload(vlo_handle *h, const float *raw, const int *offsets, int n_scans, int stride, int on_device)
{
    if (!h || !raw || !offsets || n_scans < 1 || stride < 3) return VLO_ERR_INVALID_ARG;
    if (n_scans > h->cfg.max_scans) { h->err = "n_scans exceeds max_scans"; return VLO_ERR_CAPACITY; }
    for (int s = 0; s < n_scans; s++) {
        int n = offsets[s + 1] - offsets[s];
        if (n < 0) return VLO_ERR_INVALID_ARG;
        if (n > h->cfg.max_points) { h->err = "a scan exceeds max_points"; return VLO_ERR_CAPACITY; }
    }
    cudaSetDevice(h->cfg.device);
    ScanBatchDev &sb = h->sb;
    size_t total = (size_t)(offsets[n_scans] - offsets[0]);
    // offsets staging: two pinned halves used alternately, each guarded by an event, so that the call does not have to
    // drain the stream before it returns (a device-resident batch is then enqueued without any host round trip)
    const size_t half_bytes = sizeof(int) * 2 * (size_t)h->cfg.max_scans;
    if (!h->upload_pinned) {
        VLO_CUDA(cudaMallocHost((void **)&h->upload_pinned, half_bytes * 2));
        for (int k = 0; k < 2; k++) VLO_CUDA(cudaEventCreateWithFlags(&h->upload_ev[k], cudaEventDisableTiming));
        h->upload_parity = 0; h->upload_used[0] = h->upload_used[1] = 0;
    }
    const int par = h->upload_parity; h->upload_parity ^= 1;
    if (h->upload_used[par]) VLO_CUDA(cudaEventSynchronize(h->upload_ev[par]));
    int *poff = (int *)((char *)h->upload_pinned + half_bytes * par);
    for (int s = 0; s < n_scans; s++) { poff[2 * s] = offsets[s] - offsets[0]; poff[2 * s + 1] = offsets[s + 1] - offsets[s]; }
    VLO_CUDA(cudaMemcpyAsync(sb.raw_offset, poff, sizeof(int) * 2 * (size_t)n_scans, cudaMemcpyHostToDevice, h->stream));
    VLO_CUDA(cudaEventRecord(h->upload_ev[par], h->stream));
    h->upload_used[par] = 1;
    if (on_device) {
        sb.raw = raw + (size_t)offsets[0] * stride;
    } else {
        if (total * stride > (size_t)h->cfg.max_scans * h->cfg.max_points * 4) { h->err = "raw payload exceeds staging capacity (stride > 4?)"; return VLO_ERR_CAPACITY; }
        VLO_CUDA(cudaMemcpyAsync(sb.raw_owned, raw + (size_t)offsets[0] * stride, total * stride * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        sb.raw = sb.raw_owned;
        // the caller's buffer may be pageable or reused right after the call: the copy must have left it
        VLO_CUDA(cudaStreamSynchronize(h->stream));
    }
    sb.n_scans = n_scans; sb.stride = stride; sb.scan_first = 0; sb.scan_count = n_scans;
    sb.xyz_off[0] = 0; sb.xyz_off[1] = 1; sb.xyz_off[2] = 2;
    h->online_have_last = 0;
    return VLO_OK;
}

extern "C" int vlo_scans_organise(vlo_handle *h)
{
    if (!h) return VLO_ERR_INVALID_ARG;
    if (h->sb.n_scans < 1) { h->err = "no scans uploaded"; return VLO_ERR_STATE; }
    cudaSetDevice(h->cfg.device);
    return vlo_launch_organise(h);
}

extern "C" int vlo_scans_extract(vlo_handle *h)
{
    if (!h) return VLO_ERR_INVALID_ARG;
    if (h->sb.n_scans < 1) { h->err = "no scans uploaded"; return VLO_ERR_STATE; }
    cudaSetDevice(h->cfg.device);
    h->grids_valid = 0; h->map_qmax = 0; h->lm.ds_valid = 0;
    return vlo_launch_extract(h);
}

extern "C" int vlo_scans_counts(vlo_handle *h, vlo_feature_counts *counts)
{
    if (!h || !counts) return VLO_ERR_INVALID_ARG;
    int B = h->sb.n_scans;
    std::vector<int> tmp((size_t)B * 8);
    VLO_CUDA(cudaMemcpyAsync(tmp.data(), h->sb.counts, sizeof(int) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
    int rc = vlo_synchronize(h); if (rc) return rc;
    for (int b = 0; b < B; b++) {
        counts[b].n_valid = tmp[b * 8]; counts[b].n_sharp = tmp[b * 8 + 1]; counts[b].n_less_sharp = tmp[b * 8 + 2];
        counts[b].n_flat = tmp[b * 8 + 3]; counts[b].n_less_flat = tmp[b * 8 + 4];
        h->map_qmax = std::max(h->map_qmax, tmp[b * 8 + 2] + tmp[b * 8 + 4]);     // upper bound of the down-sampled stack sizes
    }
    return VLO_OK;
}

extern "C" int vlo_scan_get_cloud(vlo_handle *h, int scan, float *cloud, int *ring_start, int *src_index)
{
    if (!h || scan < 0 || scan >= h->sb.n_scans) return VLO_ERR_INVALID_ARG;
    int rc = vlo_synchronize(h); if (rc) return rc;
    int cnt[8];
    VLO_CUDA(cudaMemcpy(cnt, h->sb.counts + scan * 8, sizeof(cnt), cudaMemcpyDeviceToHost));
    size_t N = h->cfg.max_points;
    if (cloud) VLO_CUDA(cudaMemcpy(cloud, h->sb.cloud + scan * N, sizeof(float4) * (size_t)cnt[0], cudaMemcpyDeviceToHost));
    if (src_index) VLO_CUDA(cudaMemcpy(src_index, h->sb.src_index + scan * N, sizeof(int) * (size_t)cnt[0], cudaMemcpyDeviceToHost));
    if (ring_start) VLO_CUDA(cudaMemcpy(ring_start, h->sb.ring_start + scan * (VLO_MAX_RINGS + 1), sizeof(int) * (size_t)(h->cfg.n_rings + 1), cudaMemcpyDeviceToHost));
    return VLO_OK;
}

extern "C" int vlo_scan_get_features(vlo_handle *h, int scan, int8_t *label, float *curvature, uint8_t *picked,
                                     int *sharp_idx, int *less_sharp_idx, int *flat_idx, float *less_flat,
                                     int *lsharp_ring_start, int *lflat_ring_start)
{
    if (!h || scan < 0 || scan >= h->sb.n_scans) return VLO_ERR_INVALID_ARG;
    int rc = vlo_synchronize(h); if (rc) return rc;
    ScanBatchDev &sb = h->sb;
    int cnt[8];
    VLO_CUDA(cudaMemcpy(cnt, sb.counts + scan * 8, sizeof(cnt), cudaMemcpyDeviceToHost));
    size_t N = h->cfg.max_points; int R = h->cfg.n_rings;
    if (label) VLO_CUDA(cudaMemcpy(label, sb.label + scan * N, (size_t)cnt[0], cudaMemcpyDeviceToHost));
    if (curvature) VLO_CUDA(cudaMemcpy(curvature, sb.curvature + scan * N, sizeof(float) * (size_t)cnt[0], cudaMemcpyDeviceToHost));
    if (picked) VLO_CUDA(cudaMemcpy(picked, sb.picked + scan * N, (size_t)cnt[0], cudaMemcpyDeviceToHost));
    if (sharp_idx) VLO_CUDA(cudaMemcpy(sharp_idx, sb.sharp_idx + (size_t)scan * h->cap_sharp, sizeof(int) * (size_t)cnt[1], cudaMemcpyDeviceToHost));
    if (less_sharp_idx) VLO_CUDA(cudaMemcpy(less_sharp_idx, sb.lsharp_idx + (size_t)scan * h->cap_lsharp, sizeof(int) * (size_t)cnt[2], cudaMemcpyDeviceToHost));
    if (flat_idx) VLO_CUDA(cudaMemcpy(flat_idx, sb.flat_idx + (size_t)scan * h->cap_flat, sizeof(int) * (size_t)cnt[3], cudaMemcpyDeviceToHost));
    if (lsharp_ring_start) VLO_CUDA(cudaMemcpy(lsharp_ring_start, sb.lsharp_ring_start + scan * (VLO_MAX_RINGS + 1), sizeof(int) * (size_t)(R + 1), cudaMemcpyDeviceToHost));
    if (lflat_ring_start) VLO_CUDA(cudaMemcpy(lflat_ring_start, sb.lflat_ring_start + scan * (VLO_MAX_RINGS + 1), sizeof(int) * (size_t)(R + 1), cudaMemcpyDeviceToHost));
    if (less_flat && cnt[4] > 0) VLO_CUDA(cudaMemcpy(less_flat, sb.lflat_pts + scan * N, sizeof(float4) * (size_t)cnt[4], cudaMemcpyDeviceToHost));
    return VLO_OK;
}

// ---------------------------------------------------------------------------------------------
// scan-to-scan
extern "C" int vlo_set_trace(vlo_handle *h, int enable) { if (!h) return VLO_ERR_INVALID_ARG; h->trace = enable ? 1 : 0; return VLO_OK; }

extern "C" int vlo_register_pairs(vlo_handle *h, const int *last, const int *cur, int n_pairs,
                                  const float *seeds, const float *last_transforms, vlo_result *out)
{
    if (!h || !last || !cur || !out || n_pairs < 1) return VLO_ERR_INVALID_ARG;
    if (n_pairs > h->max_pairs) { h->err = "n_pairs exceeds max_scans"; return VLO_ERR_CAPACITY; }
    for (int p = 0; p < n_pairs; p++)
        if (last[p] < 0 || last[p] >= h->sb.n_scans || cur[p] < 0 || cur[p] >= h->sb.n_scans) { h->err = "pair index outside the resident batch"; return VLO_ERR_INVALID_ARG; }
    cudaSetDevice(h->cfg.device);
    size_t need = sizeof(int) * 2 * (size_t)n_pairs + sizeof(float) * 12 * (size_t)n_pairs + sizeof(vlo_result) * (size_t)n_pairs;
    int rc = ensure_pinned(h, need); if (rc) return rc;
    char *pp = (char *)h->pinned;
    vlo_result *pres = (vlo_result *)pp; pp += sizeof(vlo_result) * (size_t)n_pairs;
    int *pl = (int *)pp; pp += sizeof(int) * (size_t)n_pairs;
    int *pc = (int *)pp; pp += sizeof(int) * (size_t)n_pairs;
    float *ps = (float *)pp; pp += sizeof(float) * 6 * (size_t)n_pairs;
    float *pt = (float *)pp;
    memcpy(pl, last, sizeof(int) * (size_t)n_pairs); memcpy(pc, cur, sizeof(int) * (size_t)n_pairs);
    VLO_CUDA(cudaMemcpyAsync(h->pair_last, pl, sizeof(int) * (size_t)n_pairs, cudaMemcpyHostToDevice, h->stream));
    VLO_CUDA(cudaMemcpyAsync(h->pair_cur, pc, sizeof(int) * (size_t)n_pairs, cudaMemcpyHostToDevice, h->stream));
    if (seeds) { memcpy(ps, seeds, sizeof(float) * 6 * (size_t)n_pairs);
        VLO_CUDA(cudaMemcpyAsync(h->pair_seed, ps, sizeof(float) * 6 * (size_t)n_pairs, cudaMemcpyHostToDevice, h->stream)); }
    if (last_transforms) { memcpy(pt, last_transforms, sizeof(float) * 6 * (size_t)n_pairs);
        VLO_CUDA(cudaMemcpyAsync(h->pair_last_T, pt, sizeof(float) * 6 * (size_t)n_pairs, cudaMemcpyHostToDevice, h->stream)); }
    rc = vlo_launch_register_pairs(h, n_pairs, seeds ? h->pair_seed : nullptr, last_transforms ? h->pair_last_T : nullptr, -1);
    if (rc) return rc;
    VLO_CUDA(cudaMemcpyAsync(pres, h->pair_result, sizeof(vlo_result) * (size_t)n_pairs, cudaMemcpyDeviceToHost, h->stream));
    rc = vlo_synchronize(h); if (rc) return rc;
    memcpy(out, pres, sizeof(vlo_result) * (size_t)n_pairs);
    for (int p = 0; p < n_pairs; p++) vlo_finish_cov_host(&out[p], &h->cfg);
    h->last_n_pairs = n_pairs;
    int soft = VLO_OK;
    for (int p = 0; p < n_pairs; p++) if (out[p].status == VLO_SOFT_TOO_FEW_CORR) soft = VLO_SOFT_TOO_FEW_CORR;
    return soft;
}

extern "C" int vlo_register_pairs_enqueue(vlo_handle *h, const int *last, const int *cur, int n_pairs, const float *seeds, vlo_result *out_pinned)
{
    if (!h || !last || !cur || !out_pinned || n_pairs < 1) return VLO_ERR_INVALID_ARG;
    if (n_pairs > h->max_pairs) { h->err = "n_pairs exceeds max_scans"; return VLO_ERR_CAPACITY; }
    for (int p = 0; p < n_pairs; p++)
        if (last[p] < 0 || last[p] >= h->sb.n_scans || cur[p] < 0 || cur[p] >= h->sb.n_scans) { h->err = "pair index outside the resident batch"; return VLO_ERR_INVALID_ARG; }
    cudaSetDevice(h->cfg.device);
    // (small copies from pageable memory are staged by the runtime before the call returns: see vlo_register_map_enqueue)
    VLO_CUDA(cudaMemcpyAsync(h->pair_last, last, sizeof(int) * (size_t)n_pairs, cudaMemcpyHostToDevice, h->stream));
    VLO_CUDA(cudaMemcpyAsync(h->pair_cur, cur, sizeof(int) * (size_t)n_pairs, cudaMemcpyHostToDevice, h->stream));
    if (seeds) VLO_CUDA(cudaMemcpyAsync(h->pair_seed, seeds, sizeof(float) * 6 * (size_t)n_pairs, cudaMemcpyHostToDevice, h->stream));
    int rc = vlo_launch_register_pairs(h, n_pairs, seeds ? h->pair_seed : nullptr, nullptr, -1);
    if (rc) return rc;
    VLO_CUDA(cudaMemcpyAsync(out_pinned, h->pair_result, sizeof(vlo_result) * (size_t)n_pairs, cudaMemcpyDeviceToHost, h->stream));
    h->last_n_pairs = n_pairs;
    return VLO_OK;
}

extern "C" int vlo_pair_get_correspondences(vlo_handle *h, int pair, int round, int *corner_idx, int *surf_idx)
{
    if (!h || pair < 0 || pair >= h->last_n_pairs || round < 0 || round >= 5) return VLO_ERR_INVALID_ARG;
    if (!h->trace) { h->err = "enable tracing with vlo_set_trace before registering"; return VLO_ERR_STATE; }
    int rc = vlo_synchronize(h); if (rc) return rc;
    int P = h->last_n_pairs;
    size_t trace_stride = (size_t)P * (h->cap_sharp * 2 + h->cap_flat * 3);
    const int *base = h->pair_trace + (size_t)round * trace_stride;
    int cur; VLO_CUDA(cudaMemcpy(&cur, h->pair_cur + pair, sizeof(int), cudaMemcpyDeviceToHost));
    int cnt[8]; VLO_CUDA(cudaMemcpy(cnt, h->sb.counts + cur * 8, sizeof(cnt), cudaMemcpyDeviceToHost));
    if (corner_idx && cnt[1] > 0) VLO_CUDA(cudaMemcpy(corner_idx, base + (size_t)pair * h->cap_sharp * 2, sizeof(int) * 2 * (size_t)cnt[1], cudaMemcpyDeviceToHost));
    if (surf_idx && cnt[3] > 0) VLO_CUDA(cudaMemcpy(surf_idx, base + (size_t)P * h->cap_sharp * 2 + (size_t)pair * h->cap_flat * 3, sizeof(int) * 3 * (size_t)cnt[3], cudaMemcpyDeviceToHost));
    return VLO_OK;
}
