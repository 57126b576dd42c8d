// Internal definitions shared by the sm_100a kernels and the C-ABI (include/vlo.h).
// Compiled with --fmad=false: every float32 expression that is compared bit-for-bit against the
// oracle must evaluate as separate IEEE mul/add (see DESIGN.md "Determinism").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/vlo.h"

#define K0_TILE 512           // raw points per CTA of the organise passes (k0_organise.cu)
#define VLO_NTERM 28          // 21 upper-tri AtA + 6 AtB + sum of squared weighted residuals
#define VLO_PI_D 3.14159265358979323846

// A kernel's dynamic shared memory as an int array.  Under the CPU emulation of tests/host/ (VLO_HOST_EMULATION) shared
// memory is ordinary static storage, one CTA at a time.
#ifdef VLO_HOST_EMULATION
#define VLO_DYN_SMEM_INT(name) static int name[16384]
#else
#define VLO_DYN_SMEM_INT(name) extern __shared__ int name[]
#endif

// One set of voxel-hash grids (see grid.cuh for the layout contract)
struct GridSet {
    float cell, inv_cell;
    int ts;              // table size per grid (power of two)
    int max_pts;
    int G;
    unsigned long long *keys;
    int *cnt;
    int *start;
    float4 *sorted;
    int *bsum;           // [G][ts / GRID_SCAN_BLOCK] block sums of the build's parallel scan
};
#define GRID_SCAN_BLOCK 4096

// where a grid build reads its points from: ring-slotted clouds (ring r's live points at
// [ring_off[r], +ring_cnt[r]) of a per-scan base array; dense index = dense_start[r] + offset) or
// plain dense arrays (ring taken from int(intensity)).
struct GridSource {
    const float4 *pts; size_t pts_stride;
    const int *ring_off; int ring_off_stride;
    const int *ring_cnt; int ring_cnt_stride;
    const int *dense_start; int dense_start_stride;
    const int *n_dense; int n_dense_stride; int n_dense_field;
    int n_rings;
    const int *grid_scan;                        // optional: grid g reads scan grid_scan[g]
    // optional sub-map filter (maintained map, k7_map.cu): point i is taken iff its cube is live and masked
    const int *cube;                             // [n] packed cube coordinates | LM_DEAD
    unsigned mask[42];                           // (2 nb + 1)^3 bits, x fastest
    int mask_lo[3], mask_side;                   // cube coordinates of mask bit (0,0,0); side = 2 nb + 1
};

#ifndef SEG_PTS
#define SEG_PTS 8          // points per fine segment of the ring-segment box index (segbox.cuh)
#endif
// ring-segment box indices of the scan-to-scan target clouds (segbox.cuh): per resident scan, cloud 0 = less sharp, 1 = less flat
struct SegSet {
    float4 *fbox[2]; float4 *mbox[2]; float4 *cbox[2]; int *perm[2]; int *seg_ring[2]; int *nseg[2];
    int4 *prange[2];          // [B][cap_lsharp] / [B][max_points] per target point: where upstream's partner loops started there break (.x backward,
                              // .y forward) and the arcs [.z, .w) of the box index in between (k3_odometry.cu)
    int max_seg[2], max_coarse[2];
};

// Device-resident state of a batch of scans (capacities from vlo_config)
struct ScanBatchDev {
    int    n_scans;
    int    scan_first, scan_count;   // range the next organise / extract launch covers
    int    stride;            // floats per raw point
    int    xyz_off[3];        // float offsets of the x, y, z fields inside a raw point (PointCloud2 fields[].offset / 4)
    const float *raw;         // device pointer (owned or borrowed)
    float *raw_owned;
    int   *raw_offset;        // [B+1] device
    // K0
    int   *first_half;        // [B]  first valid index with halfPassed condition
    float *ori_bounds;        // [B][2] startOri, endOri
    int   *tile_hist;         // [B][R][tiles]
    int8_t *ring_of; float *ori_of;   // [B][N] ring id (-1 = dropped) and raw orientation per raw point (K0 pass 1 -> pass 3)
    float4 *cloud;            // [B][N] ring-major
    int   *ring_start;        // [B][R+1]
    int   *src_index;         // [B][N]
    // K1
    int8_t *label;            // [B][N]
    float  *curvature;        // [B][N]
    uint8_t *picked;          // [B][N]
    int   *slot_sharp;        // [B][R][NR][max_sharp]   cloud indices, -1 padded
    int   *slot_lsharp;       // [B][R][NR][max_lsharp]
    int   *slot_flat;         // [B][R][NR][max_flat]
    uint8_t *slot_cnt;        // [B][R][NR][4]  (sharp, lsharp, flat, unused)
    float4 *lflat_slotted;    // [B][N]  K1 scratch: ring r's centroids at [ring_start[r], +lflat_cnt[r]); K1b packs them into lflat_pts
    int   *lflat_cnt;         // [B][R]
    // K1b dense feature packs
    int   *counts;            // [B][8]: n_valid n_sharp n_lsharp n_flat n_lflat status
    int   *sharp_idx, *lsharp_idx, *flat_idx;   // [B][cap]
    float4 *sharp_pts, *lsharp_pts, *flat_pts;  // [B][cap]
    float4 *lflat_pts;                          // [B][N] dense less-flat cloud (ring-major, ring r at [lflat_ring_start[r], ..))
    int   *lsharp_ring_start, *lflat_ring_start; // [B][R+1]
};

// Maintained map (LaserMapping's map side, k7_map.cu).  Per cloud w (0 corner, 1 surface): leaf hash
// (cube, voxel) -> point index, per-point cube tag, accumulators of the insertion in flight.
#define LM_DEAD (1 << 30)
struct LaserMapDev {
    int lts;                              // leaf hash size (power of two)
    unsigned long long *lkeys[2]; int *lval[2]; int *lfirst[2];
    int *cube[2];                         // [cap]
    int *acc[2];                          // [cap][4] sum x y z (2^-20 m), count
    int *stamp[2]; uint8_t *fresh[2];     // [cap]
    float4 *pm[2]; int *qslot[2]; int *qcube[2];   // per stack point temporaries [qcap_w]
    float4 *ins_pts[2]; int *ins_n;       // staging of vlo_map_insert's host points; ins_n[8] like counts rows
    float *ins_T;                         // [6]
    int tick;
    int cen[3];                           // laserCloudCenWidth / Height / Depth
    int *sub_n;                           // device [8]: [2] / [4] = sub-map sizes (what k5 gates on)
    // stack down-sampling (per resident scan)
    float4 *ds_pts[2]; int *ds_counts;    // [B][cap_lsharp] / [B][N]; [B][8] fields 2 / 4
    unsigned long long *ds_keys; int *ds_rec; int *ds_slot;   // [B][hts]; [B][hts][8]; [B][cap_lsharp + N]
    int ds_hoff[2], ds_hsize[2], ds_hts;
    int ds_valid;                         // down-sampled stacks of the resident batch are current
    int mode;                             // 0 no map, 1 static (vlo_map_build), 2 maintained
};

struct vlo_handle {
    vlo_config cfg;
    cudaStream_t stream;
    std::string err;
    long long launches;
    int tiles_per_scan;
    int cap_sharp, cap_lsharp, cap_flat;
    ScanBatchDev sb;
    int *status_word;          // device: sticky error bits from kernels
    // registration workspace
    int   max_pairs;
    float *pair_T;             // [P][6]
    float *pair_seed;
    int   *pair_last, *pair_cur;
    int   *pair_state;         // [P][4]: converged, iterations, is_degenerate, status
    int   *pair_cidx;          // [P][cap_sharp][2]
    int   *pair_sidx;          // [P][cap_flat][3]
    int   *pair_trace;         // first-association copy for parity tests (5 rounds)
    vlo_result *pair_result;   // device results [P]
    float *pair_last_T;        // [P][6] staging of last_transforms
    int grids_valid, trace, last_n_pairs;
    int scan_index_grid;       // the resident scans' search index: 1 voxel-hash grids (a few scans: the online tick), 0 ring-segment boxes
    // box indices of the scan-to-scan targets (per resident scan) and the voxel-hash grids of the map (0 corner, 1 surf)
    GridSet gs_corner, gs_surf;
    SegSet segs;
    GridSet gs_map[2];
    float4 *map_pts[2];
    int *map_n;                // device [8]: [2] = n_corner, [4] = n_surf (same layout as counts rows)
    int map_n_host[2];
    // mapping workspace (slot k of a vlo_register_map call)
    float *map_partials;       // [n][pcap][28] level-1 sums
    int   *map_idx5;           // [n][qcap][5]
    float *map_T; float *map_seed; int *map_state; int *map_ncorr; int *map_done; int *map_scans; vlo_result *map_result;
    int map_qmax, last_n_map;
    int coop_resident;
    // per-device launch configuration, cached per HANDLE (function attributes and occupancy are per device; a process may
    // hold handles on several devices)
    size_t k1_smem_configured, k1c_smem_configured;
    int dev_sms, k5_occ_assoc, k5_occ_lin, k3_gn_configured, k3a_ctas;
    int k0_sub, k1_sub;        // tuning (VLO_K0_SUB / VLO_K1_SUB): scans per sub-batch of K0 / K1; 0 = the whole batch in one pass
    LaserMapDev lm;
    // IMU staging (grown on demand)
    double *imu_buf; size_t imu_buf_bytes; vlo_preint *imu_out; int imu_out_cap;
    // stage profiling
    int prof_enabled; std::vector<cudaEvent_t> prof_events; std::vector<int> prof_stage; size_t prof_used;
    // pinned staging
    void *pinned; size_t pinned_bytes;
    void *upload_pinned; cudaEvent_t upload_ev[2]; int upload_parity, upload_used[2];   // vlo_scans_upload's offsets staging
    void *bag_ctx;             // whole-bag streaming: copy stream, events, staging (vlo_bag.cu), created on first use
    // online state
    int online_have_last; float online_T[6]; float online_sum[6]; float online_map_bef[6], online_map_aft[6];
    int online_slot; long long online_ticks;
};

// per-stage device timing (bench.py's roofline leg): CUDA events on the handle's stream around
// every launch group, summed per stage by vlo_get_stage_times
enum VloStage { ST_ORGANISE = 0, ST_EXTRACT, ST_COMPACT, ST_GRID_BUILD, ST_TO_END, ST_ASSOC, ST_GN, ST_MAP_KNN, ST_MAP_LIN,
                ST_MAP_ASSOC, ST_IMU, ST_STACK_DS, ST_MAP_INSERT, ST_COUNT };
void vlo_prof_begin(vlo_handle *h, int stage);
void vlo_prof_end(vlo_handle *h, int stage);
#define VLO_PROF(h, stage, stmt) do { vlo_prof_begin(h, stage); stmt; vlo_prof_end(h, stage); } while (0)

#define VLO_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return VLO_ERR_CUDA; } } while (0)

// ----------------------------------------------------------------------------------------------
// deterministic float32 elementary functions (Cephes single-precision kernels, no FMA): the same
// published algorithm the oracle restates in oracle/detmath.h.
__device__ __forceinline__ void vlo_sincosf(float x, float &s_out, float &c_out)
{
    const float two_over_pi = 0.63661977236758134308f;
    const float P1 = 1.5703125f, P2 = 4.837512969970703125e-4f, P3 = 7.54978995489188216e-8f;
    float kf = rintf(x * two_over_pi);
    int k = (int)kf;
    float r = x - kf * P1;
    r = r - kf * P2;
    r = r - kf * P3;
    float z = r * r;
    float sp = -1.9515295891e-4f * z;
    sp = sp + 8.3321608736e-3f;
    sp = sp * z;
    sp = sp - 1.6666654611e-1f;
    sp = sp * z;
    sp = sp * r;
    float sn = sp + r;
    float cp = 2.443315711809948e-5f * z;
    cp = cp - 1.388731625493765e-3f;
    cp = cp * z;
    cp = cp + 4.166664568298827e-2f;
    cp = cp * z;
    cp = cp * z;
    float hz = 0.5f * z;
    float cs = cp - hz;
    cs = cs + 1.0f;
    switch (k & 3) {
    case 0: s_out = sn;  c_out = cs;  break;
    case 1: s_out = cs;  c_out = -sn; break;
    case 2: s_out = -sn; c_out = -cs; break;
    default: s_out = -cs; c_out = sn; break;
    }
}

__device__ __forceinline__ float vlo_atanf(float xx)
{
    float x = fabsf(xx), y;
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }
    else y = 0.0f;
    float z = x * x;
    float p = 8.05374449538e-2f * z;
    p = p - 1.38776856032e-1f;
    p = p * z;
    p = p + 1.99777106478e-1f;
    p = p * z;
    p = p - 3.33329491539e-1f;
    p = p * z;
    p = p * x;
    p = p + x;
    y = y + p;
    return (xx < 0.0f) ? -y : y;
}

__device__ __forceinline__ float vlo_atan2f(float y, float x)
{
    const float PI_F = 3.14159265358979323846f, PIO2_F = 1.5707963267948966f;
    if (x == 0.0f) { if (y > 0.0f) return PIO2_F; if (y < 0.0f) return -PIO2_F; return 0.0f; }
    float a = vlo_atanf(y / x);
    if (x < 0.0f) { if (y < 0.0f) return a - PI_F; return a + PI_F; }
    return a;
}

__device__ __forceinline__ float sqdiff3(float ax, float ay, float az, float bx, float by, float bz)
{
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    return (dx * dx + dy * dy) + dz * dz;
}

// transformToStart: T = rx ry rz tx ty tz
__device__ __forceinline__ float4 vlo_to_start(const float *T, float4 p, int deskew, float inv_period)
{
    float s = deskew ? inv_period * (p.w - (float)(int)p.w) : 1.0f;
    float x = p.x - s * T[3], y = p.y - s * T[4], z = p.z - s * T[5];
    float sx, cx, sy, cy, sz, cz;
    vlo_sincosf(-s * T[0], sx, cx);
    vlo_sincosf(-s * T[1], sy, cy);
    vlo_sincosf(-s * T[2], sz, cz);
    float x0 = x; x = cz * x0 - sz * y; y = sz * x0 + cz * y;          // rotZ
    float y0 = y; y = cx * y0 - sx * z; z = sx * y0 + cx * z;          // rotX
    x0 = x;       x = cy * x0 + sy * z; z = cy * z - sy * x0;          // rotY
    return make_float4(x, y, z, p.w);
}

// pointAssociateToMap: rotZ(rz) rotX(rx) rotY(ry) then + t; trig = srx crx sry cry srz crz
__device__ __forceinline__ float4 vlo_to_map(const float *T, const float *trig, float4 pi)
{
    float x = pi.x, y = pi.y, z = pi.z;
    float sx = trig[0], cx = trig[1], sy = trig[2], cy = trig[3], sz = trig[4], cz = trig[5];
    float x0 = x; x = cz * x0 - sz * y; y = sz * x0 + cz * y;
    float y0 = y; y = cx * y0 - sx * z; z = sx * y0 + cx * z;
    x0 = x;       x = cy * x0 + sy * z; z = cy * z - sy * x0;
    return make_float4(x + T[3], y + T[4], z + T[5], pi.w);
}

// kernels' host launchers -----------------------------------------------------------------------
void vlo_finish_cov_host(vlo_result *r, const vlo_config *cfg);
int vlo_launch_organise(vlo_handle *h);
int vlo_launch_extract(vlo_handle *h);
int vlo_grid_build(vlo_handle *h, const GridSet &gs, const GridSource &src, int g_first, int n_grids, int n_slots);
int vlo_grid_knn(vlo_handle *h, const GridSet &gs, int g, const float4 *d_q, int nq, int k, float dmax, int *d_idx, float *d_d2);
int vlo_build_scan_grids(vlo_handle *h, int first, int count);
int vlo_launch_register_pairs(vlo_handle *h, int n_pairs, const float *d_seeds, const float *d_last_T, int only_grid_scan);
int vlo_launch_to_end(vlo_handle *h, const int *d_scans, const float *d_T, int n);
int vlo_launch_register_map(vlo_handle *h, const int *d_scans, int n, const float *d_seeds);
int vlo_launch_stack_ds(vlo_handle *h, int first, int count);
int vlo_lm_alloc(vlo_handle *h);
void vlo_lm_free(vlo_handle *h);
void vlo_bag_free(vlo_handle *h);
int vlo_launch_imu(vlo_handle *h, const double *d_t, const double *d_acc, const double *d_gyro, int n_samples,
                   const double *d_t0, const double *d_t1, const double *d_bias, int n_factors, vlo_preint *d_out);
