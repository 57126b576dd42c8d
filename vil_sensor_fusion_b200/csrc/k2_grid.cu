// K2  voxel-hash build (count -> scan -> scatter) for a set of grids, plus a k-NN service kernel.
// See grid.cuh for the layout and the exactness contract.
#include "grid.cuh"
#include <algorithm>

__device__ __forceinline__ bool grid_src_point(const GridSource &src, int g, int i, float4 &p, unsigned &tag)
{
    int b = src.grid_scan ? src.grid_scan[g] : g;
    if (src.ring_off) {
        const int *ro = src.ring_off + (size_t)b * src.ring_off_stride;
        if (i >= ro[src.n_rings]) return false;
        int lo = 0, hi = src.n_rings;            // ring r with ro[r] <= i < ro[r+1]
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (ro[mid] <= i) lo = mid; else hi = mid; }
        int off = i - ro[lo];
        if (off >= src.ring_cnt[(size_t)b * src.ring_cnt_stride + lo]) return false;
        p = src.pts[(size_t)b * src.pts_stride + i];
        int dense = src.dense_start[(size_t)b * src.dense_start_stride + lo] + off;
        tag = ((unsigned)lo << 24) | (unsigned)dense;
        return true;
    }
    int n = src.n_dense[(size_t)b * src.n_dense_stride + src.n_dense_field];
    if (i >= n) return false;
    if (src.cube) {                              // sub-map of the maintained map: live points of the masked cubes
        const int cv = src.cube[i];
        if (cv & LM_DEAD) return false;
        const int r0 = ((cv & 1023) - 512) - src.mask_lo[0], r1 = (((cv >> 10) & 1023) - 512) - src.mask_lo[1],
                  r2 = (((cv >> 20) & 1023) - 512) - src.mask_lo[2], sd = src.mask_side;
        if (r0 < 0 || r0 >= sd || r1 < 0 || r1 >= sd || r2 < 0 || r2 >= sd) return false;
        const int bit = (r2 * sd + r1) * sd + r0;
        if (!((src.mask[bit >> 5] >> (bit & 31)) & 1u)) return false;
    }
    p = src.pts[(size_t)b * src.pts_stride + i];
    tag = ((unsigned)(int)p.w << 24) | (unsigned)i;
    return true;
}

__global__ void __launch_bounds__(256) k2_count(GridSet gs, GridSource src, int n_slots, int g_first)
{
    int g = g_first + blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    float4 p; unsigned tag;
    if (!grid_src_point(src, g, i, p, tag)) return;
    int ix = (int)floorf(p.x * gs.inv_cell), iy = (int)floorf(p.y * gs.inv_cell), iz = (int)floorf(p.z * gs.inv_cell);
    unsigned long long key = grid_key(ix, iy, iz);
    unsigned long long *keys = gs.keys + (size_t)g * gs.ts;
    int slot = (int)(grid_hash(ix, iy, iz) & (unsigned)(gs.ts - 1));
    while (true) {
        unsigned long long old = atomicCAS(&keys[slot], GRID_EMPTY, key);
        if (old == GRID_EMPTY || old == key) break;
        slot = (slot + 1) & (gs.ts - 1);
    }
    atomicAdd(&gs.cnt[(size_t)g * gs.ts + slot], 1);
}

// exclusive scan of cnt -> start, three passes over blocks of GRID_SCAN_BLOCK slots (1024 threads x 4):
//   k2_scan_partial  block sums;  k2_scan_bsums  one CTA per grid scans them (+ start[ts] = total);
//   k2_scan_final    block-local exclusive scan + block offset
__device__ __forceinline__ int k2_block_excl_scan4(int v0, int v1, int v2, int v3, int *warp_sum, int &block_total)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mine = v0 + v1 + v2 + v3;
    int inc = mine;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sum[lane], winc = w;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += u; }
        warp_sum[lane] = winc - w;
        if (lane == 31) warp_sum[32] = winc;
    }
    __syncthreads();
    block_total = warp_sum[32];
    return warp_sum[warp] + inc - mine;
}

__global__ void __launch_bounds__(1024) k2_scan_partial(GridSet gs, int g_first)
{
    __shared__ int warp_sum[33];
    const int g = g_first + blockIdx.y, nblk = (gs.ts + GRID_SCAN_BLOCK - 1) / GRID_SCAN_BLOCK;
    const int base = blockIdx.x * GRID_SCAN_BLOCK + threadIdx.x * 4;
    int4 v = make_int4(0, 0, 0, 0);
    if (base < gs.ts) v = *(const int4 *)(gs.cnt + (size_t)g * gs.ts + base);
    int total;
    k2_block_excl_scan4(v.x, v.y, v.z, v.w, warp_sum, total);
    if (threadIdx.x == 0) gs.bsum[(size_t)g * nblk + blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k2_scan_bsums(GridSet gs, int g_first)
{
    __shared__ int warp_sum[33];
    __shared__ int carry_s;
    const int g = g_first + blockIdx.x, nblk = (gs.ts + GRID_SCAN_BLOCK - 1) / GRID_SCAN_BLOCK, tid = threadIdx.x;
    int *bs = gs.bsum + (size_t)g * nblk;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + tid;
        const int v = i < nblk ? bs[i] : 0;
        int total;
        const int ex = k2_block_excl_scan4(v, 0, 0, 0, warp_sum, total);
        const int carry = carry_s;
        if (i < nblk) bs[i] = carry + ex;
        __syncthreads();
        if (tid == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (tid == 0) gs.start[(size_t)g * (gs.ts + 1) + gs.ts] = carry_s;
}

__global__ void __launch_bounds__(1024) k2_scan_final(GridSet gs, int g_first)
{
    __shared__ int warp_sum[33];
    const int g = g_first + blockIdx.y, nblk = (gs.ts + GRID_SCAN_BLOCK - 1) / GRID_SCAN_BLOCK;
    const int base = blockIdx.x * GRID_SCAN_BLOCK + threadIdx.x * 4;
    int4 v = make_int4(0, 0, 0, 0);
    if (base < gs.ts) v = *(const int4 *)(gs.cnt + (size_t)g * gs.ts + base);
    int total;
    const int ex = k2_block_excl_scan4(v.x, v.y, v.z, v.w, warp_sum, total) + gs.bsum[(size_t)g * nblk + blockIdx.x];
    if (base < gs.ts) {
        int *st = gs.start + (size_t)g * (gs.ts + 1) + base;      // rows of ts + 1 ints: not 16-byte aligned in general
        st[0] = ex; st[1] = ex + v.x; st[2] = ex + v.x + v.y; st[3] = ex + v.x + v.y + v.z;
    }
}

__global__ void __launch_bounds__(256) k2_scatter(GridSet gs, GridSource src, int n_slots, int g_first)
{
    int g = g_first + blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    float4 p; unsigned tag;
    if (!grid_src_point(src, g, i, p, tag)) return;
    int ix = (int)floorf(p.x * gs.inv_cell), iy = (int)floorf(p.y * gs.inv_cell), iz = (int)floorf(p.z * gs.inv_cell);
    int slot = grid_find(gs, g, ix, iy, iz);
    int pos = gs.start[(size_t)g * (gs.ts + 1) + slot] + atomicSub(&gs.cnt[(size_t)g * gs.ts + slot], 1) - 1;
    gs.sorted[(size_t)g * gs.max_pts + pos] = make_float4(p.x, p.y, p.z, __uint_as_float(tag));
}

int vlo_grid_build(vlo_handle *h, const GridSet &gs, const GridSource &src, int g_first, int n_grids, int n_slots)
{
    if (n_grids <= 0) return VLO_OK;
    vlo_prof_begin(h, ST_GRID_BUILD);
    VLO_CUDA(cudaMemsetAsync(gs.keys + (size_t)g_first * gs.ts, 0xFF, sizeof(unsigned long long) * (size_t)n_grids * gs.ts, h->stream));
    VLO_CUDA(cudaMemsetAsync(gs.cnt + (size_t)g_first * gs.ts, 0, sizeof(int) * (size_t)n_grids * gs.ts, h->stream));
    dim3 grid((n_slots + 255) / 256, n_grids);
    k2_count<<<grid, 256, 0, h->stream>>>(gs, src, n_slots, g_first);
    const int nblk = (gs.ts + GRID_SCAN_BLOCK - 1) / GRID_SCAN_BLOCK;
    k2_scan_partial<<<dim3(nblk, n_grids), 1024, 0, h->stream>>>(gs, g_first);
    k2_scan_bsums<<<n_grids, 1024, 0, h->stream>>>(gs, g_first);
    k2_scan_final<<<dim3(nblk, n_grids), 1024, 0, h->stream>>>(gs, g_first);
    k2_scatter<<<grid, 256, 0, h->stream>>>(gs, src, n_slots, g_first);
    vlo_prof_end(h, ST_GRID_BUILD);
    h->launches += 5;
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}

// k-NN service (parity tests / tooling): one warp per query
template <int K>
__global__ void __launch_bounds__(256) k2_knn(GridSet gs, int g, const float4 *q, int nq, float dmax, int *idx, float *d2)
{
    __shared__ int scratch[8][GRID_SCRATCH_INTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nq; w += n_warps) {
        float4 p = q[w];
        TopK<K> best;
        grid_search<K>(gs, g, p.x, p.y, p.z, dmax, FilterAll(), best, lane, scratch[warp]);
        if (lane == 0) {
            #pragma unroll
            for (int i = 0; i < K; i++) {
                bool ok = best.tag[i] != GRID_NOTAG;
                idx[(size_t)w * K + i] = ok ? (int)(best.tag[i] & 0xFFFFFFu) : -1;
                d2[(size_t)w * K + i] = ok ? __uint_as_float(best.d[i]) : __int_as_float(0x7f800000);
            }
        }
    }
}

int vlo_grid_knn(vlo_handle *h, const GridSet &gs, int g, const float4 *d_q, int nq, int k, float dmax, int *d_idx, float *d_d2)
{
    int blocks = std::min((nq * 32 + 255) / 256, 148 * 8);
    if (nq <= 0) return VLO_OK;
    if (k == 1) k2_knn<1><<<blocks, 256, 0, h->stream>>>(gs, g, d_q, nq, dmax, d_idx, d_d2);
    else if (k == 5) k2_knn<5><<<blocks, 256, 0, h->stream>>>(gs, g, d_q, nq, dmax, d_idx, d_d2);
    else { h->err = "k must be 1 or 5"; return VLO_ERR_INVALID_ARG; }
    h->launches += 1;
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}
