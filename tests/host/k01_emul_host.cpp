// TEST INFRASTRUCTURE -- K0 (organise) and K1 (feature extraction) of the product run on the CPU under the SIMT emulator
// (cuda_emul.h): the kernels of csrc/k0_organise.cu and csrc/k1_extract.cu AND their real host launchers
// (vlo_launch_organise / vlo_launch_extract; only the <<< >>> syntax is rewritten, gen_emul.py), on host buffers.
// tests/test_host_kernels.py compares the results with the oracle bit for bit.
#define VLO_HOST_EMULATION
#include "cuda_emul.h"
// the launchers' few runtime calls
#define cudaGetLastError() (cudaSuccess)
#define cudaFuncSetAttribute(...) (cudaSuccess)
#define cudaGetErrorString(e) "emulated"
struct vlo_handle;
void vlo_prof_begin(vlo_handle *, int) {}
void vlo_prof_end(vlo_handle *, int) {}

#include "k0_organise.emul.cpp"
#include "k1_extract.emul.cpp"

template <class T> static T *zalloc(size_t n) { return (T *)calloc(n ? n : 1, sizeof(T)); }

extern "C" int emu_organise_extract(const vlo_config *cfg, const float *raw, int n, int stride,
                                    float *cloud, int *ring_start, int *src_index, int8_t *label, float *curvature, uint8_t *picked,
                                    int *counts, int *sharp_idx, int *lsharp_idx, int *flat_idx, float *less_flat,
                                    int *lsharp_ring_start, int *lflat_ring_start)
{
    vlo_handle *h = new vlo_handle();
    h->cfg = *cfg; h->stream = nullptr; h->launches = 0;
    const int B = 1, N = cfg->max_points, R = cfg->n_rings, NR = cfg->feature_regions;
    if (n > N) return -3;
    h->tiles_per_scan = (N + K0_TILE - 1) / K0_TILE;
    h->cap_sharp = R * NR * std::max(cfg->max_corner_sharp, 1);
    h->cap_lsharp = R * NR * std::max(cfg->max_corner_less_sharp, 1);
    h->cap_flat = R * NR * std::max(cfg->max_surface_flat, 1);
    h->status_word = zalloc<int>(1);
    ScanBatchDev &sb = h->sb;
    sb.n_scans = 1; sb.scan_first = 0; sb.scan_count = 1; sb.stride = stride; sb.xyz_off[0] = 0; sb.xyz_off[1] = 1; sb.xyz_off[2] = 2;
    sb.raw = raw; sb.raw_owned = nullptr;
    sb.raw_offset = zalloc<int>(2 * B); sb.raw_offset[0] = 0; sb.raw_offset[1] = n;
    sb.first_half = zalloc<int>(B); sb.ori_bounds = zalloc<float>(2 * B);
    sb.tile_hist = zalloc<int>((size_t)B * R * h->tiles_per_scan);
    sb.ring_of = zalloc<int8_t>((size_t)B * N); sb.ori_of = zalloc<float>((size_t)B * N);
    sb.cloud = zalloc<float4>((size_t)B * N); sb.ring_start = zalloc<int>(B * (VLO_MAX_RINGS + 1)); sb.src_index = zalloc<int>((size_t)B * N);
    sb.label = zalloc<int8_t>((size_t)B * N); sb.curvature = zalloc<float>((size_t)B * N); sb.picked = zalloc<uint8_t>((size_t)B * N);
    sb.slot_sharp = zalloc<int>((size_t)B * h->cap_sharp); sb.slot_lsharp = zalloc<int>((size_t)B * h->cap_lsharp); sb.slot_flat = zalloc<int>((size_t)B * h->cap_flat);
    sb.slot_cnt = zalloc<uint8_t>((size_t)B * R * NR * 4);
    sb.lflat_slotted = zalloc<float4>((size_t)B * N); sb.lflat_cnt = zalloc<int>(B * R);
    sb.counts = zalloc<int>(B * 8);
    sb.sharp_idx = zalloc<int>((size_t)B * h->cap_sharp); sb.lsharp_idx = zalloc<int>((size_t)B * h->cap_lsharp); sb.flat_idx = zalloc<int>((size_t)B * h->cap_flat);
    sb.sharp_pts = zalloc<float4>((size_t)B * h->cap_sharp); sb.lsharp_pts = zalloc<float4>((size_t)B * h->cap_lsharp); sb.flat_pts = zalloc<float4>((size_t)B * h->cap_flat);
    sb.lflat_pts = zalloc<float4>((size_t)B * N);
    sb.lsharp_ring_start = zalloc<int>(B * (VLO_MAX_RINGS + 1)); sb.lflat_ring_start = zalloc<int>(B * (VLO_MAX_RINGS + 1));
    int rc = vlo_launch_organise(h);
    if (rc == 0) rc = vlo_launch_extract(h);
    const int nv = sb.counts[0];
    memcpy(cloud, sb.cloud, sizeof(float4) * (size_t)nv); memcpy(src_index, sb.src_index, sizeof(int) * (size_t)nv);
    memcpy(ring_start, sb.ring_start, sizeof(int) * (size_t)(R + 1));
    memcpy(label, sb.label, (size_t)nv); memcpy(curvature, sb.curvature, sizeof(float) * (size_t)nv); memcpy(picked, sb.picked, (size_t)nv);
    memcpy(counts, sb.counts, sizeof(int) * 8);
    memcpy(sharp_idx, sb.sharp_idx, sizeof(int) * (size_t)sb.counts[1]); memcpy(lsharp_idx, sb.lsharp_idx, sizeof(int) * (size_t)sb.counts[2]);
    memcpy(flat_idx, sb.flat_idx, sizeof(int) * (size_t)sb.counts[3]); memcpy(less_flat, sb.lflat_pts, sizeof(float4) * (size_t)sb.counts[4]);
    memcpy(lsharp_ring_start, sb.lsharp_ring_start, sizeof(int) * (size_t)(R + 1)); memcpy(lflat_ring_start, sb.lflat_ring_start, sizeof(int) * (size_t)(R + 1));
    return rc | (h->status_word[0] << 8);
}
