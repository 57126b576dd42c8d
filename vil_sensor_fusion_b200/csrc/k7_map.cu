// K7  LaserMapping's map side: stack down-sampling, cube-window sub-map, insertion + voxel re-filtering.
// Replaces the map half of BasicLaserMapping::process of the `loam` nodelet laserMapping
// (gtsam_fusion/launch/loam.launch:47-52; knobs loam_params.yaml:47-52: cornerFilterSize, surfaceFilterSize,
// mapCubeSize, mapDimensionsInCubes, mapStartLocationInCubes, numNeighborSubmapCubes); SURVEY.md Appendix A.8,
// oracle/laser_map.c is the frozen operation order (choices M1-M5 there).
//
// B200 design: the map is NOT an array of per-cube point clouds with a kd-tree rebuilt per tick.  It is one flat
// array of voxel centroids per cloud (index = order of voxel creation) with a packed cube tag per point and a leaf
// hash (cube, voxel) -> index.  Per tick
//   k7_ds_bin / k7_ds_emit      VoxelGrid of the sweep's corner / surface stacks: global-memory hash per scan, integer
//                               (order-free) centroid sums, output in order of first appearance (ordered block scan);
//                               the leaders clean their hash slots, so the tables never need a memset
//   vlo_grid_build (k2)         the search grids are REBUILT from the flat arrays through the cube mask of the
//                               FOV-valid neighbourhood: 1M points = 16 MB read + 16 MB written, a few tens of
//                               microseconds at HBM speed -- cheaper and simpler than mutating a spatial index
//   k5 (k5_mapping.cu)          optimisation against the sub-map
//   k7_ins_probe / _assign / _accum / _final   insertion: find-or-create in the leaf hash, new voxels numbered in stack
//                               order (ordered scan, deterministic), integer centroid sums, one finisher per voxel
#include "grid.cuh"
#include <algorithm>
#include <cstring>
#include <cmath>

#define LM_EMPTY 0xFFFFFFFFFFFFFFFFull
#define LM_VOX_BIAS (1 << 18)
#define LM_QF 1048576.0f

// ------------------------------------------------------------------------------------------------ stack VoxelGrid
struct StackDsParams {
    const float4 *src[2]; size_t stride[2];     // less-sharp [B][cap_lsharp], less-flat [B][N]
    const int *counts;                          // [B][8], fields 2 / 4
    float leaf[2];
    float4 *dst[2]; int *ds_counts;
    unsigned long long *keys; int *rec; int *slot_of;
    int hoff[2], hsize[2], hts, qstride, qoff1;
    int scan_first;
    int wmask;                                  // bit w set: cloud w is filtered here (leaf > 0)
};

// table size actually used for a stack of n points: the smallest power of two >= 2 n (at most the allocated one), so
// that the slots a sweep touches stay L2-resident instead of being spread over a worst-case-sized table
__device__ __forceinline__ int k7_ds_table_size(int n, int alloc)
{
    int hs = 1024;
    while (hs < 2 * n) hs <<= 1;
    return min(hs, alloc);
}

// record of a hash slot: sx sy sz sw cnt first pad pad (one 32-byte sector)
// Consecutive points of a ring fall into the same 0.2 / 0.4 m voxel more often than not.  The lanes of a warp hold
// consecutive points, so a voxel shows up as a RUN of lanes: run heads come from one ballot, the run's integer sums from
// a five-step segmented shuffle reduction, and the head lane alone issues the atomics (count = run length, first =
// the head's index).  A voxel that re-appears later in the warp is simply a second run with its own atomics -- the
// sums are order-free.  (r01c ncu: match_any + REDUX on per-group masks were 35 % of this kernel's samples, and 87 % of
// the 131 072 CTAs of the capacity-sized grid had nothing to do; the grid now strides over the points.)
#define K7_BIN_CTAS 24
__global__ void __launch_bounds__(256) k7_ds_bin(StackDsParams p)
{
    const int b = p.scan_first + blockIdx.y, w = blockIdx.z, lane = threadIdx.x & 31;
    if (!((p.wmask >> w) & 1)) return;
    const int n = p.counts[b * 8 + (w ? 4 : 2)];
    const float leaf = p.leaf[w], inv = 1.0f / leaf;
    const int hs = k7_ds_table_size(n, p.hsize[w]);
    unsigned long long *keys = p.keys + (size_t)b * p.hts + p.hoff[w];
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {     // CTA-uniform bounds
        const int i = base + threadIdx.x;
        const bool valid = i < n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) v = p.src[w][(size_t)b * p.stride[w] + i];
        const int ix = (int)floorf(v.x * inv), iy = (int)floorf(v.y * inv), iz = (int)floorf(v.z * inv);
        // lanes past the end get private keys (bit 63 is never set in a grid key)
        const unsigned long long key = valid ? grid_key(ix, iy, iz) : (0x8000000000000000ull | (unsigned long long)lane);
        const float ox = (float)ix * leaf, oy = (float)iy * leaf, oz = (float)iz * leaf;
        int qx = (int)rintf((v.x - ox) * LM_QF), qy = (int)rintf((v.y - oy) * LM_QF), qz = (int)rintf((v.z - oz) * LM_QF);
        int qw = (int)rintf((v.w - (float)(int)v.w) * LM_QF);
        const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != key);
        const unsigned later = heads & ~((2u << lane) - 1u);                   // heads strictly after this lane
        const int run_end = later ? (__ffs(later) - 2) : 31;                     // last lane of this lane's run
        const int leader = 31 - __clz(heads & ((2u << lane) - 1u));             // head lane of this lane's run
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int ax = __shfl_down_sync(0xffffffffu, qx, d), ay = __shfl_down_sync(0xffffffffu, qy, d);
            const int az = __shfl_down_sync(0xffffffffu, qz, d), aw = __shfl_down_sync(0xffffffffu, qw, d);
            if (lane + d <= run_end) { qx += ax; qy += ay; qz += az; qw += aw; }
        }
        int slot = 0;
        if (valid && lane == leader) {
            slot = (int)(grid_hash(ix, iy, iz) & (unsigned)(hs - 1));
            while (true) {
                unsigned long long old = atomicCAS(&keys[slot], LM_EMPTY, key);
                if (old == LM_EMPTY || old == key) break;
                slot = (slot + 1) & (hs - 1);
            }
            int *rec = p.rec + ((size_t)b * p.hts + p.hoff[w] + slot) * 8;
            atomicAdd(&rec[0], qx); atomicAdd(&rec[1], qy); atomicAdd(&rec[2], qz); atomicAdd(&rec[3], qw);
            atomicAdd(&rec[4], run_end - lane + 1);
            atomicMin(&rec[5], i);
        }
        slot = __shfl_sync(0xffffffffu, slot, leader);
        if (valid) p.slot_of[(size_t)b * p.qstride + (w ? p.qoff1 : 0) + i] = slot;
    }
}

// ordered block scan of one flag per thread (1024 threads); returns the exclusive rank, `total` = block total
__device__ __forceinline__ int k7_block_rank(bool flag, int *warp_buf, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    __syncthreads();                                   // warp_buf may still be read from the previous round
    if (lane == 0) warp_buf[warp] = __popc(bal);
    __syncthreads();
    int off = 0; total = 0;
    #pragma unroll 8
    for (int wv = 0; wv < 32; wv++) { int c = warp_buf[wv]; if (wv < warp) off += c; total += c; }
    return off + __popc(bal & ((1u << lane) - 1u));
}

// Four points per thread and round: the four (slot, leader-index) load chains of a thread are independent, so the
// L2 latency of the hash records overlaps; leaders are ranked in index order = (sub-chunk j, warp, lane).
#define K7_EMIT_PER 4
__global__ void __launch_bounds__(1024) k7_ds_emit(StackDsParams p)
{
    __shared__ int s_cnt[K7_EMIT_PER * 32 + 1];
    const int b = p.scan_first + blockIdx.x, w = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!((p.wmask >> w) & 1)) return;
    const int n = p.counts[b * 8 + (w ? 4 : 2)];
    const float leaf = p.leaf[w], inv = 1.0f / leaf;
    unsigned long long *keys = p.keys + (size_t)b * p.hts + p.hoff[w];
    int *recs = p.rec + ((size_t)b * p.hts + p.hoff[w]) * 8;
    const int *slot_of = p.slot_of + (size_t)b * p.qstride + (w ? p.qoff1 : 0);
    float4 *dst = p.dst[w] + (size_t)b * p.stride[w];
    int carry = 0;
    for (int base = 0; base < n; base += 1024 * K7_EMIT_PER) {
        int slot[K7_EMIT_PER]; unsigned bal[K7_EMIT_PER];
        #pragma unroll
        for (int j = 0; j < K7_EMIT_PER; j++) { const int i = base + j * 1024 + tid; slot[j] = i < n ? slot_of[i] : -1; }
        #pragma unroll
        for (int j = 0; j < K7_EMIT_PER; j++) {
            const int i = base + j * 1024 + tid;
            const bool lead = slot[j] >= 0 && recs[slot[j] * 8 + 5] == i;
            bal[j] = __ballot_sync(0xffffffffu, lead);
        }
        if (lane == 0) {
            #pragma unroll
            for (int j = 0; j < K7_EMIT_PER; j++) s_cnt[j * 32 + warp] = __popc(bal[j]);
        }
        __syncthreads();
        if (warp == 0) {
            int loc[K7_EMIT_PER], sum = 0;
            #pragma unroll
            for (int e = 0; e < K7_EMIT_PER; e++) { loc[e] = s_cnt[lane * K7_EMIT_PER + e]; sum += loc[e]; }
            int inc = sum;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
            int run = inc - sum;
            #pragma unroll
            for (int e = 0; e < K7_EMIT_PER; e++) { s_cnt[lane * K7_EMIT_PER + e] = run; run += loc[e]; }
            if (lane == 31) s_cnt[K7_EMIT_PER * 32] = inc;
        }
        __syncthreads();
        #pragma unroll
        for (int j = 0; j < K7_EMIT_PER; j++) {
            if (!((bal[j] >> lane) & 1u)) continue;
            const int i = base + j * 1024 + tid;
            const int pos = carry + s_cnt[j * 32 + warp] + __popc(bal[j] & ((1u << lane) - 1u));
            int *rec = recs + slot[j] * 8;
            const float4 v = p.src[w][(size_t)b * p.stride[w] + i];
            const int ix = (int)floorf(v.x * inv), iy = (int)floorf(v.y * inv), iz = (int)floorf(v.z * inv);
            const float ox = (float)ix * leaf, oy = (float)iy * leaf, oz = (float)iz * leaf;
            const int4 sums = *(const int4 *)rec;
            const float c = (float)rec[4], q = 1.0f / 1048576.0f;
            float4 o;
            o.x = ox + ((float)sums.x / c) * q;
            o.y = oy + ((float)sums.y / c) * q;
            o.z = oz + ((float)sums.z / c) * q;
            o.w = (float)(int)v.w + ((float)sums.w / c) * q;
            dst[pos] = o;
            // the leader owns the slot: leave it clean for the next sweep (others only compare rec[5] with their index)
            *(int4 *)rec = make_int4(0, 0, 0, 0); rec[4] = 0; rec[5] = 0x7fffffff;
            keys[slot[j]] = LM_EMPTY;
        }
        carry += s_cnt[K7_EMIT_PER * 32];
        // end of the round: s_cnt is rewritten and the slots the leaders have just cleaned are compared (rec[5] == i) by the
        // next round only after every thread is here (ThreadSanitizer over the CPU emulation flagged the barrier-less
        // read of a slot another thread was still cleaning; either value compared unequal, but it was a race)
        __syncthreads();
    }
    if (tid == 0) p.ds_counts[b * 8 + (w ? 4 : 2)] = carry;
}

__global__ void k7_ds_init(unsigned long long *keys, int *rec, size_t n_slots)
{
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += (size_t)gridDim.x * blockDim.x) {
        keys[s] = LM_EMPTY;
        int *r = rec + s * 8;
        r[0] = 0; r[1] = 0; r[2] = 0; r[3] = 0; r[4] = 0; r[5] = 0x7fffffff; r[6] = 0; r[7] = 0;
    }
}

int vlo_launch_stack_ds(vlo_handle *h, int first, int count)
{
    LaserMapDev &lm = h->lm; ScanBatchDev &sb = h->sb; const vlo_config &c = h->cfg;
    if (count <= 0) return VLO_OK;
    StackDsParams p;
    p.src[0] = sb.lsharp_pts; p.stride[0] = (size_t)h->cap_lsharp; p.src[1] = sb.lflat_pts; p.stride[1] = (size_t)c.max_points;
    p.counts = sb.counts; p.leaf[0] = c.corner_filter_size; p.leaf[1] = c.surface_filter_size;
    p.dst[0] = lm.ds_pts[0]; p.dst[1] = lm.ds_pts[1]; p.ds_counts = lm.ds_counts;
    p.keys = lm.ds_keys; p.rec = lm.ds_rec; p.slot_of = lm.ds_slot;
    p.hoff[0] = lm.ds_hoff[0]; p.hoff[1] = lm.ds_hoff[1]; p.hsize[0] = lm.ds_hsize[0]; p.hsize[1] = lm.ds_hsize[1]; p.hts = lm.ds_hts;
    p.qstride = h->cap_lsharp + c.max_points; p.qoff1 = h->cap_lsharp; p.scan_first = first;
    p.wmask = (p.leaf[0] > 0.0f ? 1 : 0) | (p.leaf[1] > 0.0f ? 2 : 0);
    vlo_prof_begin(h, ST_STACK_DS);
    // a cloud whose filter is off (leaf <= 0) is copied through unchanged
    for (int w = 0; w < 2; w++) {
        if ((p.wmask >> w) & 1) continue;
        const size_t st = p.stride[w];
        VLO_CUDA(cudaMemcpyAsync(p.dst[w] + (size_t)first * st, p.src[w] + (size_t)first * st, sizeof(float4) * st * (size_t)count, cudaMemcpyDeviceToDevice, h->stream));
        VLO_CUDA(cudaMemcpy2DAsync(lm.ds_counts + first * 8 + (w ? 4 : 2), sizeof(int) * 8, sb.counts + first * 8 + (w ? 4 : 2), sizeof(int) * 8,
                                   sizeof(int), (size_t)count, cudaMemcpyDeviceToDevice, h->stream));
    }
    if (p.wmask) {
        const int qmax = std::max(h->cap_lsharp, c.max_points);
        k7_ds_bin<<<dim3(std::min((qmax + 255) / 256, K7_BIN_CTAS), count, 2), 256, 0, h->stream>>>(p);
        k7_ds_emit<<<dim3(count, 2), 1024, 0, h->stream>>>(p);
        h->launches += 2;
    }
    vlo_prof_end(h, ST_STACK_DS);
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}

// ------------------------------------------------------------------------------------------------ map insertion
struct MapInsParams {
    const float4 *pts[2]; const int *n_ptr[2];     // stack clouds (sensor frame) and their device counts
    const float *T;                                // device pose rx ry rz tx ty tz
    float leaf[2], size, half; int cen[3], dims[3];
    float4 *mpts[2]; int *mcube[2]; int *map_n; int cap;
    unsigned long long *lkeys[2]; int *lval[2]; int *lfirst[2]; int lts;
    int *acc[2]; int *stamp[2]; uint8_t *fresh[2];
    float4 *pm[2]; int *qslot[2]; int *qcube[2];
    int tick; int *status_word;
};

// upstream: int((x + size/2) / size) [+ cen]; if (x + size/2 < 0) --
__device__ __forceinline__ int lm_cube_coord(float x, float half, float size)
{
    const float v = x + half;
    int c = (int)(v / size);
    if (v < 0.0f) c--;
    return c;
}

__device__ __forceinline__ unsigned long long lm_mix64(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

// find-or-claim the leaf-hash slot of every stack point; the lowest stack index per slot becomes its leader
__global__ void __launch_bounds__(256) k7_ins_probe(MapInsParams p)
{
    const int w = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *p.n_ptr[w];
    if (i >= n) return;
    float T[6], trig[6];
    #pragma unroll
    for (int a = 0; a < 6; a++) T[a] = p.T[a];
    vlo_sincosf(T[0], trig[0], trig[1]); vlo_sincosf(T[1], trig[2], trig[3]); vlo_sincosf(T[2], trig[4], trig[5]);
    const float4 pm = vlo_to_map(T, trig, p.pts[w][i]);
    p.pm[w][i] = pm;
    const float leaf = p.leaf[w], inv = 1.0f / leaf;
    const float pc[3] = { pm.x, pm.y, pm.z };
    int cube[3]; bool ok = true;
    unsigned long long key = 0;
    #pragma unroll
    for (int a = 0; a < 3; a++) {
        cube[a] = lm_cube_coord(pc[a], p.half, p.size);
        const int rel = cube[a] + p.cen[a];
        if (rel < 0 || rel >= p.dims[a] || cube[a] < -512 || cube[a] > 511) ok = false;
        const int iv = (int)floorf(pc[a] * inv);
        if (iv < -LM_VOX_BIAS || iv >= LM_VOX_BIAS) ok = false;
        const int co = lm_cube_coord((float)iv * leaf, p.half, p.size);
        key = (key << 21) | ((unsigned long long)(unsigned)(iv + LM_VOX_BIAS) << 2) | (unsigned long long)((cube[a] - co) & 3);
    }
    if (!ok) { p.qslot[w][i] = -1; return; }
    unsigned long long *keys = p.lkeys[w];
    int slot = (int)(lm_mix64(key) & (unsigned long long)(p.lts - 1));
    // bounded probing: keys of voxels dropped for capacity stay in the table, so a full map can fill it up
    int probes = 0;
    while (true) {
        unsigned long long old = atomicCAS(&keys[slot], LM_EMPTY, key);
        if (old == LM_EMPTY || old == key) break;
        slot = (slot + 1) & (p.lts - 1);
        if (++probes >= 1024) { atomicOr(p.status_word, 2); p.qslot[w][i] = -1; return; }
    }
    atomicMin(&p.lfirst[w][slot], i);
    p.qslot[w][i] = slot;
    p.qcube[w][i] = (cube[0] + 512) | ((cube[1] + 512) << 10) | ((cube[2] + 512) << 20);
}

// new voxels are numbered in stack order (ordered scan over the leaders): one CTA per cloud
__global__ void __launch_bounds__(1024) k7_ins_assign(MapInsParams p)
{
    __shared__ int warp_buf[32];
    const int w = blockIdx.x, tid = threadIdx.x;
    const int n = *p.n_ptr[w];
    const int field = w ? 4 : 2;
    const int n_old = p.map_n[field];
    int carry = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        int slot = -1; bool lead = false, isnew = false;
        if (i < n) {
            slot = p.qslot[w][i];
            lead = slot >= 0 && p.lfirst[w][slot] == i;
            isnew = lead && p.lval[w][slot] < 0;
        }
        int total;
        const int rank = k7_block_rank(isnew, warp_buf, total);
        if (isnew) {
            const int id = n_old + carry + rank;
            if (id < p.cap) { p.lval[w][slot] = id; p.mcube[w][id] = p.qcube[w][i]; p.fresh[w][id] = 1; }
            else atomicOr(p.status_word, 2);
        } else if (lead) {
            const int id = p.lval[w][slot];
            if (p.mcube[w][id] & LM_DEAD) { p.mcube[w][id] &= ~LM_DEAD; p.fresh[w][id] = 1; }   // evicted voxel comes back empty
        }
        if (lead) p.lfirst[w][slot] = 0x7fffffff;
        carry += total;
        // the next round compares lfirst of slots whose leaders have just reset it: order the two (ThreadSanitizer over the
        // CPU emulation; either value compared unequal, but the accesses were unordered)
        __syncthreads();
    }
    if (tid == 0) p.map_n[field] = min(p.cap, n_old + carry);
}

__global__ void __launch_bounds__(256) k7_ins_accum(MapInsParams p)
{
    const int w = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *p.n_ptr[w];
    if (i >= n) return;
    const int slot = p.qslot[w][i];
    if (slot < 0) return;
    const int id = p.lval[w][slot];
    if (id < 0) return;
    const float4 pm = p.pm[w][i];
    const float leaf = p.leaf[w], inv = 1.0f / leaf;
    const float ox = (float)(int)floorf(pm.x * inv) * leaf, oy = (float)(int)floorf(pm.y * inv) * leaf, oz = (float)(int)floorf(pm.z * inv) * leaf;
    int *acc = p.acc[w] + (size_t)id * 4;
    atomicAdd(&acc[0], (int)rintf((pm.x - ox) * LM_QF));
    atomicAdd(&acc[1], (int)rintf((pm.y - oy) * LM_QF));
    atomicAdd(&acc[2], (int)rintf((pm.z - oz) * LM_QF));
    atomicAdd(&acc[3], 1);
}

// VoxelGrid of a touched voxel (M3): (old centroid + new points) / (1 + k), or new points / k for a fresh voxel
__global__ void __launch_bounds__(256) k7_ins_final(MapInsParams p)
{
    const int w = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *p.n_ptr[w];
    if (i >= n) return;
    const int slot = p.qslot[w][i];
    if (slot < 0) return;
    const int id = p.lval[w][slot];
    if (id < 0) return;
    if (atomicExch(&p.stamp[w][id], p.tick) == p.tick) return;      // somebody else finishes this voxel
    const float4 pm = p.pm[w][i];
    const float leaf = p.leaf[w], inv = 1.0f / leaf;
    const float o[3] = { (float)(int)floorf(pm.x * inv) * leaf, (float)(int)floorf(pm.y * inv) * leaf, (float)(int)floorf(pm.z * inv) * leaf };
    int *acc = p.acc[w] + (size_t)id * 4;
    const bool fresh = p.fresh[w][id] != 0;
    const float4 old = p.mpts[w][id];
    const float oc[3] = { old.x, old.y, old.z };
    const int cnt = acc[3] + (fresh ? 0 : 1);
    float out[3];
    #pragma unroll
    for (int a = 0; a < 3; a++) {
        int s = acc[a];
        if (!fresh) s += (int)rintf((oc[a] - o[a]) * LM_QF);
        out[a] = o[a] + ((float)s / (float)cnt) * (1.0f / 1048576.0f);
    }
    p.mpts[w][id] = make_float4(out[0], out[1], out[2], 0.0f);
    acc[0] = 0; acc[1] = 0; acc[2] = 0; acc[3] = 0;
    p.fresh[w][id] = 0;
}

// cubes shifted out of the window are dropped (M5)
__global__ void __launch_bounds__(256) k7_evict(int *cube, const int *map_n, int field, int c0, int c1, int c2, int d0, int d1, int d2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= map_n[field]) return;
    const int v = cube[i];
    const int r0 = ((v & 1023) - 512) + c0, r1 = (((v >> 10) & 1023) - 512) + c1, r2 = (((v >> 20) & 1023) - 512) + c2;
    if (r0 < 0 || r0 >= d0 || r1 < 0 || r1 >= d1 || r2 < 0 || r2 >= d2) cube[i] = v | LM_DEAD;
}

__global__ void k7_sub_counts(GridSet g0, GridSet g1, int *sub_n)
{
    if (threadIdx.x == 0) { sub_n[2] = g0.start[g0.ts]; sub_n[4] = g1.start[g1.ts]; }
}

__global__ void k7_lm_init(unsigned long long *keys, int *val, int *first, size_t n)
{
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (size_t)gridDim.x * blockDim.x) {
        keys[s] = LM_EMPTY; val[s] = -1; first[s] = 0x7fffffff;
    }
}

// ------------------------------------------------------------------------------------------------ host side
template <typename T> static cudaError_t lm_dalloc(T **p, size_t n) { return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)); }
#define LM_ALLOC(ptr, n) do { cudaError_t e_ = lm_dalloc(&(ptr), (n)); if (e_ != cudaSuccess) { \
    h->err = std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_); return VLO_ERR_CUDA; } } while (0)

int vlo_lm_alloc(vlo_handle *h)
{
    LaserMapDev &lm = h->lm; const vlo_config &c = h->cfg;
    const int B = c.max_scans, N = c.max_points, cap = c.max_map_points;
    // stack down-sampling
    lm.ds_hsize[0] = 1024; while (lm.ds_hsize[0] < 2 * h->cap_lsharp) lm.ds_hsize[0] <<= 1;
    lm.ds_hsize[1] = 1024; while (lm.ds_hsize[1] < 2 * N) lm.ds_hsize[1] <<= 1;
    lm.ds_hoff[0] = 0; lm.ds_hoff[1] = lm.ds_hsize[0]; lm.ds_hts = lm.ds_hsize[0] + lm.ds_hsize[1];
    LM_ALLOC(lm.ds_pts[0], (size_t)B * h->cap_lsharp); LM_ALLOC(lm.ds_pts[1], (size_t)B * N);
    LM_ALLOC(lm.ds_counts, (size_t)B * 8);
    LM_ALLOC(lm.ds_keys, (size_t)B * lm.ds_hts); LM_ALLOC(lm.ds_rec, (size_t)B * lm.ds_hts * 8);
    LM_ALLOC(lm.ds_slot, (size_t)B * (h->cap_lsharp + N));
    cudaMemsetAsync(lm.ds_counts, 0, sizeof(int) * (size_t)B * 8, h->stream);
    k7_ds_init<<<148 * 8, 256, 0, h->stream>>>(lm.ds_keys, lm.ds_rec, (size_t)B * lm.ds_hts);
    // maintained map
    lm.lts = 1024; while (lm.lts < 2 * cap) lm.lts <<= 1;
    const int qcap[2] = { std::max(h->cap_lsharp, 1), std::max(N, 1) };
    for (int w = 0; w < 2; w++) {
        LM_ALLOC(lm.lkeys[w], (size_t)lm.lts); LM_ALLOC(lm.lval[w], (size_t)lm.lts); LM_ALLOC(lm.lfirst[w], (size_t)lm.lts);
        LM_ALLOC(lm.cube[w], (size_t)cap); LM_ALLOC(lm.acc[w], (size_t)cap * 4); LM_ALLOC(lm.stamp[w], (size_t)cap); LM_ALLOC(lm.fresh[w], (size_t)cap);
        LM_ALLOC(lm.pm[w], (size_t)qcap[w]); LM_ALLOC(lm.qslot[w], (size_t)qcap[w]); LM_ALLOC(lm.qcube[w], (size_t)qcap[w]);
        LM_ALLOC(lm.ins_pts[w], (size_t)qcap[w]);
    }
    LM_ALLOC(lm.ins_n, 8); LM_ALLOC(lm.ins_T, 6); LM_ALLOC(lm.sub_n, 8);
    cudaMemsetAsync(lm.sub_n, 0, sizeof(int) * 8, h->stream);
    lm.tick = 0; lm.mode = 0; lm.ds_valid = 0;
    return VLO_OK;
}

void vlo_lm_free(vlo_handle *h)
{
    LaserMapDev &lm = h->lm;
    void *ptrs[] = { lm.ds_pts[0], lm.ds_pts[1], lm.ds_counts, lm.ds_keys, lm.ds_rec, lm.ds_slot, lm.ins_n, lm.ins_T, lm.sub_n };
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int w = 0; w < 2; w++) {
        void *q[] = { lm.lkeys[w], lm.lval[w], lm.lfirst[w], lm.cube[w], lm.acc[w], lm.stamp[w], lm.fresh[w], lm.pm[w], lm.qslot[w], lm.qcube[w], lm.ins_pts[w] };
        for (void *p : q) if (p) cudaFree(p);
    }
    memset(&lm, 0, sizeof(lm));
}

static int lm_reset_device(vlo_handle *h)
{
    LaserMapDev &lm = h->lm; const vlo_config &c = h->cfg;
    for (int w = 0; w < 2; w++) {
        k7_lm_init<<<148 * 4, 256, 0, h->stream>>>(lm.lkeys[w], lm.lval[w], lm.lfirst[w], (size_t)lm.lts);
        VLO_CUDA(cudaMemsetAsync(lm.acc[w], 0, sizeof(int) * 4 * (size_t)c.max_map_points, h->stream));
        VLO_CUDA(cudaMemsetAsync(lm.stamp[w], 0, sizeof(int) * (size_t)c.max_map_points, h->stream));
        VLO_CUDA(cudaMemsetAsync(lm.fresh[w], 0, (size_t)c.max_map_points, h->stream));
    }
    VLO_CUDA(cudaMemsetAsync(h->map_n, 0, sizeof(int) * 8, h->stream));
    VLO_CUDA(cudaMemsetAsync(lm.sub_n, 0, sizeof(int) * 8, h->stream));
    h->map_n_host[0] = h->map_n_host[1] = 0;
    for (int a = 0; a < 3; a++) lm.cen[a] = c.map_start_cubes[a];
    lm.tick = 0; lm.mode = 2;
    h->launches += 2;
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}

static int lm_check(vlo_handle *h)
{
    const vlo_config &c = h->cfg;
    if (c.max_map_points <= 0) { h->err = "handle created with max_map_points = 0"; return VLO_ERR_STATE; }
    if (!(c.corner_filter_size > 0.f) || !(c.surface_filter_size > 0.f) || !(c.map_cube_size > 0.f) || c.n_neighbor_cubes < 0 || c.n_neighbor_cubes > 5 ||
        c.map_dims[0] < 7 || c.map_dims[1] < 7 || c.map_dims[2] < 7 || c.map_dims[0] > 1023 || c.map_dims[1] > 1023 || c.map_dims[2] > 1023) {
        h->err = "maintained map needs cornerFilterSize, surfaceFilterSize, mapCubeSize > 0, numNeighborSubmapCubes <= 5, 7 <= mapDimensionsInCubes <= 1023";
        return VLO_ERR_INVALID_ARG;
    }
    return VLO_OK;
}

extern "C" int vlo_map_reset(vlo_handle *h)
{
    if (!h) return VLO_ERR_INVALID_ARG;
    int rc = lm_check(h); if (rc) return rc;
    cudaSetDevice(h->cfg.device);
    rc = lm_reset_device(h); if (rc) return rc;
    return vlo_synchronize(h);
}

// insertion of device-resident stack clouds with a device pose
static int lm_launch_insert(vlo_handle *h, const float4 *pts0, const int *n0, const float4 *pts1, const int *n1, const float *d_T, int nmax0, int nmax1)
{
    LaserMapDev &lm = h->lm; const vlo_config &c = h->cfg;
    MapInsParams p;
    p.pts[0] = pts0; p.pts[1] = pts1; p.n_ptr[0] = n0; p.n_ptr[1] = n1; p.T = d_T;
    p.leaf[0] = c.corner_filter_size; p.leaf[1] = c.surface_filter_size; p.size = c.map_cube_size; p.half = c.map_cube_size / 2.0f;
    for (int a = 0; a < 3; a++) { p.cen[a] = lm.cen[a]; p.dims[a] = c.map_dims[a]; }
    p.map_n = h->map_n; p.cap = c.max_map_points; p.lts = lm.lts;
    for (int w = 0; w < 2; w++) {
        p.mpts[w] = h->map_pts[w]; p.mcube[w] = lm.cube[w]; p.lkeys[w] = lm.lkeys[w]; p.lval[w] = lm.lval[w]; p.lfirst[w] = lm.lfirst[w];
        p.acc[w] = lm.acc[w]; p.stamp[w] = lm.stamp[w]; p.fresh[w] = lm.fresh[w]; p.pm[w] = lm.pm[w]; p.qslot[w] = lm.qslot[w]; p.qcube[w] = lm.qcube[w];
    }
    p.tick = ++lm.tick; p.status_word = h->status_word;
    const int nmax = std::max(std::max(nmax0, nmax1), 1);
    dim3 g((nmax + 255) / 256, 2);
    vlo_prof_begin(h, ST_MAP_INSERT);
    k7_ins_probe<<<g, 256, 0, h->stream>>>(p);
    k7_ins_assign<<<2, 1024, 0, h->stream>>>(p);
    k7_ins_accum<<<g, 256, 0, h->stream>>>(p);
    k7_ins_final<<<g, 256, 0, h->stream>>>(p);
    vlo_prof_end(h, ST_MAP_INSERT);
    h->launches += 4;
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}

extern "C" int vlo_map_insert(vlo_handle *h, const float *corner, int n_corner, const float *surf, int n_surf, const float *pose6)
{
    if (!h || n_corner < 0 || n_surf < 0 || (n_corner > 0 && !corner) || (n_surf > 0 && !surf) || !pose6) return VLO_ERR_INVALID_ARG;
    int rc = lm_check(h); if (rc) return rc;
    cudaSetDevice(h->cfg.device);
    LaserMapDev &lm = h->lm;
    if (lm.mode != 2) { rc = lm_reset_device(h); if (rc) return rc; }
    // the whole cloud is one insertion step (all points of a voxel are averaged together, M3): clouds larger than
    // the per-sweep staging get temporaries of their own for this call
    const int cap[2] = { std::max(h->cap_lsharp, 1), std::max(h->cfg.max_points, 1) };
    const int nn[2] = { n_corner, n_surf };
    const float *src[2] = { corner, surf };
    float4 *keep_pm[2], *keep_ins[2]; int *keep_qs[2], *keep_qc[2];
    bool own[2] = { false, false };
    for (int w = 0; w < 2; w++) {
        keep_pm[w] = lm.pm[w]; keep_ins[w] = lm.ins_pts[w]; keep_qs[w] = lm.qslot[w]; keep_qc[w] = lm.qcube[w];
        if (nn[w] > cap[w]) {
            own[w] = true;
            if (cudaMalloc((void **)&lm.pm[w], sizeof(float4) * (size_t)nn[w]) != cudaSuccess || cudaMalloc((void **)&lm.ins_pts[w], sizeof(float4) * (size_t)nn[w]) != cudaSuccess ||
                cudaMalloc((void **)&lm.qslot[w], sizeof(int) * (size_t)nn[w]) != cudaSuccess || cudaMalloc((void **)&lm.qcube[w], sizeof(int) * (size_t)nn[w]) != cudaSuccess) {
                h->err = "cudaMalloc of the insertion temporaries failed"; rc = VLO_ERR_CUDA;
            }
        }
    }
    int cnt[8] = { 0, 0, n_corner, 0, n_surf, 0, 0, 0 };
    if (rc == VLO_OK && cudaMemcpyAsync(lm.ins_T, pose6, sizeof(float) * 6, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) rc = VLO_ERR_CUDA;
    if (rc == VLO_OK && cudaMemcpyAsync(lm.ins_n, cnt, sizeof(cnt), cudaMemcpyHostToDevice, h->stream) != cudaSuccess) rc = VLO_ERR_CUDA;
    for (int w = 0; w < 2 && rc == VLO_OK; w++)
        if (nn[w] > 0 && cudaMemcpyAsync(lm.ins_pts[w], src[w], sizeof(float4) * (size_t)nn[w], cudaMemcpyHostToDevice, h->stream) != cudaSuccess) rc = VLO_ERR_CUDA;
    if (rc == VLO_OK) rc = lm_launch_insert(h, lm.ins_pts[0], lm.ins_n + 2, lm.ins_pts[1], lm.ins_n + 4, lm.ins_T, n_corner, n_surf);
    cudaStreamSynchronize(h->stream);                    // pageable sources and the temporaries are done with
    for (int w = 0; w < 2; w++) {
        if (own[w]) { cudaFree(lm.pm[w]); cudaFree(lm.ins_pts[w]); cudaFree(lm.qslot[w]); cudaFree(lm.qcube[w]); }
        lm.pm[w] = keep_pm[w]; lm.ins_pts[w] = keep_ins[w]; lm.qslot[w] = keep_qs[w]; lm.qcube[w] = keep_qc[w];
    }
    if (rc) return rc;
    int mn[8];
    VLO_CUDA(cudaMemcpy(mn, h->map_n, sizeof(mn), cudaMemcpyDeviceToHost));
    h->map_n_host[0] = mn[2]; h->map_n_host[1] = mn[4];
    return vlo_synchronize(h);
}

extern "C" int vlo_map_size(vlo_handle *h, int *n_corner, int *n_surf)
{
    if (!h) return VLO_ERR_INVALID_ARG;
    int rc = vlo_synchronize(h); if (rc) return rc;
    int mn[8];
    VLO_CUDA(cudaMemcpy(mn, h->map_n, sizeof(mn), cudaMemcpyDeviceToHost));
    if (n_corner) *n_corner = mn[2];
    if (n_surf) *n_surf = mn[4];
    return VLO_OK;
}

extern "C" int vlo_map_get_points(vlo_handle *h, int which, float *xyzi, int *cube)
{
    if (!h || which < 0 || which > 1) return VLO_ERR_INVALID_ARG;
    if (h->cfg.max_map_points <= 0) { h->err = "handle created with max_map_points = 0"; return VLO_ERR_STATE; }
    int rc = vlo_synchronize(h); if (rc) return rc;
    int mn[8];
    VLO_CUDA(cudaMemcpy(mn, h->map_n, sizeof(mn), cudaMemcpyDeviceToHost));
    const int n = mn[which ? 4 : 2];
    if (xyzi && n > 0) VLO_CUDA(cudaMemcpy(xyzi, h->map_pts[which], sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost));
    if (cube && n > 0) {
        if (h->lm.mode == 2) VLO_CUDA(cudaMemcpy(cube, h->lm.cube[which], sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
        else memset(cube, 0, sizeof(int) * (size_t)n);
    }
    return VLO_OK;
}

extern "C" int vlo_scan_get_stack(vlo_handle *h, int scan, float *corner, int *n_corner, float *surf, int *n_surf)
{
    if (!h || scan < 0 || scan >= h->sb.n_scans) return VLO_ERR_INVALID_ARG;
    if (h->cfg.max_map_points <= 0) { h->err = "handle created with max_map_points = 0"; return VLO_ERR_STATE; }
    cudaSetDevice(h->cfg.device);
    if (!h->lm.ds_valid) { int rc = vlo_launch_stack_ds(h, 0, h->sb.n_scans); if (rc) return rc; h->lm.ds_valid = 1; }
    int rc = vlo_synchronize(h); if (rc) return rc;
    int cnt[8];
    VLO_CUDA(cudaMemcpy(cnt, h->lm.ds_counts + scan * 8, sizeof(cnt), cudaMemcpyDeviceToHost));
    if (n_corner) *n_corner = cnt[2];
    if (n_surf) *n_surf = cnt[4];
    if (corner && cnt[2] > 0) VLO_CUDA(cudaMemcpy(corner, h->lm.ds_pts[0] + (size_t)scan * h->cap_lsharp, sizeof(float4) * (size_t)cnt[2], cudaMemcpyDeviceToHost));
    if (surf && cnt[4] > 0) VLO_CUDA(cudaMemcpy(surf, h->lm.ds_pts[1] + (size_t)scan * h->cfg.max_points, sizeof(float4) * (size_t)cnt[4], cudaMemcpyDeviceToHost));
    return VLO_OK;
}

extern "C" int vlo_scans_stack_counts(vlo_handle *h, int *n_corner, int *n_surf)
{
    if (!h || !n_corner || !n_surf) return VLO_ERR_INVALID_ARG;
    if (h->cfg.max_map_points <= 0) { h->err = "handle created with max_map_points = 0"; return VLO_ERR_STATE; }
    cudaSetDevice(h->cfg.device);
    if (!h->lm.ds_valid) { int rc = vlo_launch_stack_ds(h, 0, h->sb.n_scans); if (rc) return rc; h->lm.ds_valid = 1; }
    const int B = h->sb.n_scans;
    std::vector<int> tmp((size_t)B * 8);
    VLO_CUDA(cudaMemcpyAsync(tmp.data(), h->lm.ds_counts, sizeof(int) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
    int rc = vlo_synchronize(h); if (rc) return rc;
    for (int b = 0; b < B; b++) { n_corner[b] = tmp[b * 8 + 2]; n_surf[b] = tmp[b * 8 + 4]; }
    return VLO_OK;
}

// steps 1-2 of the tick on the host (a few thousand flops): window shift around `T`, FOV mask of the neighbourhood.
// Same float32 operations as oracle/laser_map.c orc_lmap_select.
static float lm_sincos_host(float x, float *c_out);
static int lm_cube_coord_host(float x, float half, float size) { float v = x + half; int c = (int)(v / size); if (v < 0.0f) c--; return c; }

static void lm_select_host(vlo_handle *h, const float *T, GridSource &src, bool &shifted)
{
    LaserMapDev &lm = h->lm; const vlo_config &c = h->cfg;
    const float size = c.map_cube_size, half = size / 2.0f;
    const int nb = c.n_neighbor_cubes, side = 2 * nb + 1;
    int centre[3];
    shifted = false;
    for (int a = 0; a < 3; a++) {
        int ca = lm_cube_coord_host(T[3 + a], half, size);
        int cc = ca + lm.cen[a];
        while (cc < 3) { cc++; lm.cen[a]++; shifted = true; }
        while (cc >= c.map_dims[a] - 3) { cc--; lm.cen[a]--; shifted = true; }
        centre[a] = ca;
    }
    // pointOnYAxis = pointAssociateToMap((0, 10, 0)) with the deterministic sin/cos of the device path
    float sx, cx, sy, cy, sz, cz;
    sx = lm_sincos_host(T[0], &cx); sy = lm_sincos_host(T[1], &cy); sz = lm_sincos_host(T[2], &cz);
    float x = 0.0f, y = 10.0f, z = 0.0f;
    float x0 = x; x = cz * x0 - sz * y; y = sz * x0 + cz * y;
    float y0 = y; y = cx * y0 - sx * z; z = sx * y0 + cx * z;
    x0 = x;       x = cy * x0 + sy * z; z = cy * z - sy * x0;
    const float yax[3] = { x + T[3], y + T[4], z + T[5] };
    const float s3 = 10.0f * sqrtf(3.0f);
    memset(src.mask, 0, sizeof(src.mask));
    for (int dk = 0; dk < side; dk++) for (int dj = 0; dj < side; dj++) for (int di = 0; di < side; di++) {
        const int cb[3] = { centre[0] - nb + di, centre[1] - nb + dj, centre[2] - nb + dk };
        bool ok = true;
        for (int a = 0; a < 3; a++) { int rel = cb[a] + lm.cen[a]; if (rel < 0 || rel >= c.map_dims[a]) ok = false; }
        if (!ok) continue;
        const float ccx = size * (float)cb[0], ccy = size * (float)cb[1], ccz = size * (float)cb[2];
        bool fov = false;
        for (int ii = -1; ii <= 1 && !fov; ii += 2) for (int jj = -1; jj <= 1 && !fov; jj += 2) for (int kk = -1; kk <= 1 && !fov; kk += 2) {
            const float px = ccx + half * (float)ii, py = ccy + half * (float)jj, pz = ccz + half * (float)kk;
            const float ax = T[3] - px, ay = T[4] - py, az = T[5] - pz;
            const float bx = yax[0] - px, by = yax[1] - py, bz = yax[2] - pz;
            const float s1 = (ax * ax + ay * ay) + az * az;
            const float s2 = (bx * bx + by * by) + bz * bz;
            const float r = s3 * sqrtf(s1);
            const float check1 = ((100.0f + s1) - s2) - r;
            const float check2 = ((100.0f + s1) - s2) + r;
            if (check1 < 0.0f && check2 > 0.0f) fov = true;
        }
        if (fov) { const int bit = (dk * side + dj) * side + di; src.mask[bit >> 5] |= 1u << (bit & 31); }
    }
    for (int a = 0; a < 3; a++) src.mask_lo[a] = centre[a] - nb;
    src.mask_side = side;
}

// Cephes single-precision sin/cos, no FMA: the host twin of vlo_sincosf (vlo_internal.cuh) / oracle/detmath.h
static float lm_sincos_host(float xin, float *c_out)
{
    volatile float x = xin;          // volatile: every operation rounds to float32, no contraction by the host compiler
    const float two_over_pi = 0.63661977236758134308f;
    const float P1 = 1.5703125f, P2 = 4.837512969970703125e-4f, P3 = 7.54978995489188216e-8f;
    volatile float kf = rintf(x * two_over_pi);
    int k = (int)kf;
    volatile float r = x - kf * P1;
    r = r - kf * P2;
    r = r - kf * P3;
    volatile float z = r * r;
    volatile float sp = -1.9515295891e-4f * z;
    sp = sp + 8.3321608736e-3f; sp = sp * z; sp = sp - 1.6666654611e-1f; sp = sp * z; sp = sp * r;
    volatile float sn = sp + r;
    volatile float cp = 2.443315711809948e-5f * z;
    cp = cp - 1.388731625493765e-3f; cp = cp * z; cp = cp + 4.166664568298827e-2f; cp = cp * z; cp = cp * z;
    volatile float hz = 0.5f * z;
    volatile float cs = cp - hz;
    cs = cs + 1.0f;
    float s_out;
    switch (k & 3) {
    case 0: s_out = sn;  *c_out = cs;  break;
    case 1: s_out = cs;  *c_out = -sn; break;
    case 2: s_out = -sn; *c_out = -cs; break;
    default: s_out = -cs; *c_out = sn; break;
    }
    return s_out;
}

static int ensure_pinned_lm(vlo_handle *h)
{
    const size_t bytes = (size_t)1 << 12;
    if (bytes <= h->pinned_bytes) return VLO_OK;
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr; h->pinned_bytes = 0;
    VLO_CUDA(cudaMallocHost(&h->pinned, bytes));
    h->pinned_bytes = bytes;
    return VLO_OK;
}

extern "C" int vlo_map_process(vlo_handle *h, int scan, const float *seed6, vlo_result *out, int *info)
{
    if (!h || !seed6 || !out || scan < 0 || scan >= h->sb.n_scans) return VLO_ERR_INVALID_ARG;
    int rc = lm_check(h); if (rc) return rc;
    cudaSetDevice(h->cfg.device);
    LaserMapDev &lm = h->lm; const vlo_config &c = h->cfg;
    if (lm.mode != 2) { rc = lm_reset_device(h); if (rc) return rc; }
    if (!lm.ds_valid) { rc = vlo_launch_stack_ds(h, 0, h->sb.n_scans); if (rc) return rc; lm.ds_valid = 1; }
    // 1-2: window + sub-map grids
    bool shifted = false;
    GridSource sel = {};
    lm_select_host(h, seed6, sel, shifted);
    if (shifted) {
        for (int v = 0; v < 2; v++)
            if (h->map_n_host[v] > 0)
                k7_evict<<<(h->map_n_host[v] + 255) / 256, 256, 0, h->stream>>>(lm.cube[v], h->map_n, v == 0 ? 2 : 4, lm.cen[0], lm.cen[1], lm.cen[2],
                                                                             c.map_dims[0], c.map_dims[1], c.map_dims[2]);
        h->launches += 2;
    }
    for (int w = 0; w < 2; w++) {
        GridSource src = sel;
        src.pts = h->map_pts[w]; src.n_dense = h->map_n; src.n_dense_field = w == 0 ? 2 : 4; src.n_rings = 1;
        src.cube = lm.cube[w];
        rc = vlo_grid_build(h, h->gs_map[w], src, 0, 1, std::max(h->map_n_host[w], 1)); if (rc) return rc;
    }
    k7_sub_counts<<<1, 32, 0, h->stream>>>(h->gs_map[0], h->gs_map[1], lm.sub_n);
    h->launches += 1;
    // 3-4: registration of the down-sampled stacks against the sub-map
    rc = ensure_pinned_lm(h); if (rc) return rc;
    char *pp = (char *)h->pinned;
    vlo_result *pres = (vlo_result *)pp; pp += sizeof(vlo_result);
    int *pinfo = (int *)pp; pp += sizeof(int) * 32;
    float *pseed = (float *)pp; pp += sizeof(float) * 8;
    int *pscan = (int *)pp;
    memcpy(pseed, seed6, sizeof(float) * 6); *pscan = scan;
    VLO_CUDA(cudaMemcpyAsync(h->map_scans, pscan, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    VLO_CUDA(cudaMemcpyAsync(h->map_seed, pseed, sizeof(float) * 6, cudaMemcpyHostToDevice, h->stream));
    h->map_qmax = 0;
    rc = vlo_launch_register_map(h, h->map_scans, 1, h->map_seed); if (rc) return rc;
    // 5: insertion with the optimised pose (the seed when the optimisation was skipped)
    rc = lm_launch_insert(h, lm.ds_pts[0] + (size_t)scan * h->cap_lsharp, lm.ds_counts + scan * 8 + 2,
                          lm.ds_pts[1] + (size_t)scan * c.max_points, lm.ds_counts + scan * 8 + 4, h->map_T, h->cap_lsharp, c.max_points);
    if (rc) return rc;
    VLO_CUDA(cudaMemcpyAsync(pres, h->map_result, sizeof(vlo_result), cudaMemcpyDeviceToHost, h->stream));
    VLO_CUDA(cudaMemcpyAsync(pinfo, lm.ds_counts + scan * 8, sizeof(int) * 8, cudaMemcpyDeviceToHost, h->stream));
    VLO_CUDA(cudaMemcpyAsync(pinfo + 8, lm.sub_n, sizeof(int) * 8, cudaMemcpyDeviceToHost, h->stream));
    VLO_CUDA(cudaMemcpyAsync(pinfo + 16, h->map_n, sizeof(int) * 8, cudaMemcpyDeviceToHost, h->stream));
    rc = vlo_synchronize(h);
    if (rc && rc != VLO_ERR_CAPACITY) return rc;
    // a capacity report (map full: new voxels were dropped) leaves the registration result and the map valid: the record
    // is delivered and the code returned beside it
    h->map_n_host[0] = pinfo[18]; h->map_n_host[1] = pinfo[20];
    *out = *pres;
    vlo_finish_cov_host(out, &h->cfg);
    h->last_n_map = 1;
    if (info) { info[0] = pinfo[2]; info[1] = pinfo[4]; info[2] = pinfo[10]; info[3] = pinfo[12]; info[4] = pinfo[18]; info[5] = pinfo[20]; }
    if (rc) return rc;
    return out->status == VLO_SOFT_TOO_FEW_CORR ? VLO_SOFT_TOO_FEW_CORR : VLO_OK;
}
