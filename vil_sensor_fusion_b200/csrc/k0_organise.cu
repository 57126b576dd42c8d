// K0  organise: raw PointCloud2 payload (AoS float32, ROS axes) -> ring-major float4 cloud in the
// LOAM frame with intensity = ring + relTime.  Replaces MultiScanRegistration::process of the
// `loam` nodelet the reference launches (gtsam_fusion/launch/loam.launch:33-38; knobs
// loam_params.yaml:3,22).  SURVEY.md Appendix A.1 is the algorithm; the sequential `halfPassed`
// flag becomes "index of the first valid point whose branch-A orientation passes pi" (a min
// reduction) and the per-ring push_back becomes a stable 3-kernel counting sort by ring.
#include "vlo_internal.cuh"
#include <cmath>
#include <algorithm>


struct K0Params {
    const float *raw; const int *raw_offset; int stride; int scan_first;   // raw_offset: [b] = {begin, count}
    int xo, yo, zo;   // float offsets of the ROS x, y, z fields inside a point
    int n_rings; float lower_deg, factor, scan_period;
    int N;            // capacity per scan
    int tiles;        // tiles per scan (capacity)
    int rotate; float R[9];   // rotateInputCloud / inputCloudRotation: R = Rz(yaw) Ry(pitch) Rx(roll), ROS frame
    int ring_field;   // ring field of the point: float index (ring_type 0, FLOAT32) or byte offset (1 UINT16, 2 UINT8); -1 = ring from the vertical angle
    int ring_type;
};

// LOAM-frame x / y / z (x = ROS y, y = ROS z, z = ROS x) of a raw point, rotated first when rotateInputCloud is set
// (oracle: ros_point; products summed as (r0 x + r1 y) + r2 z)
__device__ __forceinline__ void k0_point(const K0Params &p, const float *q, float &x, float &y, float &z)
{
    float rx = q[p.xo], ry = q[p.yo], rz = q[p.zo];
    if (p.rotate) {
        const float ax = (p.R[0] * rx + p.R[1] * ry) + p.R[2] * rz;
        const float ay = (p.R[3] * rx + p.R[4] * ry) + p.R[5] * rz;
        const float az = (p.R[6] * rx + p.R[7] * ry) + p.R[8] * rz;
        rx = ax; ry = ay; rz = az;
    }
    x = ry; y = rz; z = rx;
}

// `rf`: the point's ring field when p.ring_field >= 0 (ring ids as delivered by the driver), unused otherwise
__device__ __forceinline__ int k0_ring(const K0Params &p, float x, float y, float z, float rf)
{
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) return -1;
    if ((x * x + y * y) + z * z < 0.0001f) return -1;
    if (p.ring_field >= 0) return (rf >= 0.0f && rf < (float)p.n_rings) ? (int)rf : -1;      // also rejects NaN
    float angle = vlo_atanf(y / sqrtf(x * x + z * z));
    float a180 = angle * 180.0f;
    double v = ((double)a180 / VLO_PI_D - (double)p.lower_deg) * (double)p.factor + 0.5;
    int id = (int)v;
    if (id >= p.n_rings || id < 0) return -1;
    return id;
}

__global__ void k0_bounds(K0Params p, float *ori_bounds, int *first_half, int n_scans)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_scans) return;
    b += p.scan_first;
    int o0 = p.raw_offset[2 * b], o1 = o0 + p.raw_offset[2 * b + 1];
    first_half[b] = 0x7fffffff;
    if (o1 <= o0) { ori_bounds[2 * b] = 0.f; ori_bounds[2 * b + 1] = 0.f; return; }
    const float *f = p.raw + (size_t)o0 * p.stride, *l = p.raw + (size_t)(o1 - 1) * p.stride;
    float fx, fy, fz, lx, ly, lz;                 // LOAM x = ROS y, LOAM z = ROS x
    k0_point(p, f, fx, fy, fz);
    k0_point(p, l, lx, ly, lz);
    float startOri = -vlo_atan2f(fx, fz);
    float endOri = -vlo_atan2f(lx, lz) + 2.0f * (float)VLO_PI_D;
    if ((double)(endOri - startOri) > 3 * VLO_PI_D) endOri = (float)((double)endOri - 2 * VLO_PI_D);
    else if ((double)(endOri - startOri) < VLO_PI_D) endOri = (float)((double)endOri + 2 * VLO_PI_D);
    ori_bounds[2 * b] = startOri; ori_bounds[2 * b + 1] = endOri;
}

// pass 1: per-tile ring histogram + first index passing the half-sweep test
__global__ void __launch_bounds__(K0_TILE) k0_classify(K0Params p, const float *ori_bounds, int *first_half, int *tile_hist,
                                                        int8_t *ring_of, float *ori_of)
{
    __shared__ int hist[VLO_MAX_RINGS];
    __shared__ int s_first;
    int b = p.scan_first + blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    int o0 = p.raw_offset[2 * b], n = p.raw_offset[2 * b + 1];
    if (tile * K0_TILE >= n) {   // still must zero the histogram column for the scan kernel
        if (tid < p.n_rings) tile_hist[((size_t)b * p.n_rings + tid) * p.tiles + tile] = 0;
        return;
    }
    if (tid < p.n_rings) hist[tid] = 0;
    if (tid == 0) s_first = 0x7fffffff;
    __syncthreads();
    int i = tile * K0_TILE + tid;
    if (i < n) {
        const float *q = p.raw + (size_t)(o0 + i) * p.stride;
        float x, y, z;
        k0_point(p, q, x, y, z);
        float rf = 0.0f;
        if (p.ring_field >= 0) {
            if (p.ring_type == 0) rf = q[p.ring_field];
            else {
                const unsigned char *bp = reinterpret_cast<const unsigned char *>(q) + p.ring_field;
                rf = p.ring_type == 1 ? (float)(bp[0] | (bp[1] << 8)) : (float)bp[0];
            }
        }
        int ring = k0_ring(p, x, y, z, rf);
        ring_of[(size_t)b * p.N + i] = (int8_t)ring;            // the scatter pass reuses ring and raw orientation
        if (ring >= 0) {
            atomicAdd(&hist[ring], 1);
            float startOri = ori_bounds[2 * b];
            float ori = -vlo_atan2f(x, z);
            ori_of[(size_t)b * p.N + i] = ori;
            if ((double)ori < (double)startOri - VLO_PI_D / 2) ori = (float)((double)ori + 2 * VLO_PI_D);
            else if ((double)ori > (double)startOri + VLO_PI_D * 3 / 2) ori = (float)((double)ori - 2 * VLO_PI_D);
            if ((double)(ori - startOri) > VLO_PI_D) atomicMin(&s_first, i);
        }
    }
    __syncthreads();
    if (tid < p.n_rings) tile_hist[((size_t)b * p.n_rings + tid) * p.tiles + tile] = hist[tid];
    if (tid == 0 && s_first != 0x7fffffff) atomicMin(&first_half[b], s_first);
}

// pass 2: per (scan, ring) exclusive scan over tiles; ring_start
__global__ void __launch_bounds__(256) k0_scan(K0Params p, int *tile_hist, int *ring_start, int *counts)
{
    __shared__ int ring_total[VLO_MAX_RINGS];
    int b = p.scan_first + blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    int n = p.raw_offset[2 * b + 1];
    int tiles = (n + K0_TILE - 1) / K0_TILE;
    for (int r = warp; r < p.n_rings; r += nw) {
        int *row = tile_hist + ((size_t)b * p.n_rings + r) * p.tiles;
        int carry = 0;
        for (int t0 = 0; t0 < tiles; t0 += 32) {
            int t = t0 + lane;
            int v = (t < tiles) ? row[t] : 0;
            int inc = v;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
            if (t < tiles) row[t] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) ring_total[r] = carry;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int r = 0; r < p.n_rings; r++) { ring_start[b * (VLO_MAX_RINGS + 1) + r] = acc; acc += ring_total[r]; }
        for (int r = p.n_rings; r <= VLO_MAX_RINGS; r++) ring_start[b * (VLO_MAX_RINGS + 1) + r] = acc;
        counts[b * 8 + 0] = acc;
    }
}

// pass 3: stable scatter into ring-major order, rel-time with the final half-sweep rule.  The tile is first sorted by
// ring in shared memory, so that each ring's run of the tile leaves as one contiguous, coalesced burst.
// 128 threads per 512-point tile, FOUR points per thread: warp w owns the tile's 32-point chunks 4 w .. 4 w + 3 (their loads are
// issued together), 16 CTAs per SM.  (Round 1 ran 512 threads with one point each: 4 CTAs per SM, every one of them waiting at
// five 16-warp barriers -- ncu: 11.8 cycles per issue at the barrier, 25 % of the DRAM throughput.)
#define K0S_THREADS 128
#define K0S_U (K0_TILE / K0S_THREADS)            // points per thread = chunks per warp
#define K0S_CHUNKS (K0_TILE / 32)
__global__ void __launch_bounds__(K0S_THREADS) k0_scatter(K0Params p, const float *ori_bounds, const int *first_half,
                                                          const int *tile_hist, const int *ring_start,
                                                          const int8_t *ring_of, const float *ori_of, float4 *cloud, int *src_index)
{
    __shared__ int chunk_cnt[K0S_CHUNKS][VLO_MAX_RINGS];
    __shared__ int run_start[VLO_MAX_RINGS + 1];     // sorted position where ring r's run of this tile starts
    __shared__ int run_dst[VLO_MAX_RINGS];           // global position (within the scan) of that run
    __shared__ float4 sorted[K0_TILE];
    __shared__ int sorted_src[K0_TILE];
    __shared__ int8_t sorted_ring[K0_TILE];
    const int b = p.scan_first + blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int o0 = p.raw_offset[2 * b], n = p.raw_offset[2 * b + 1];
    if (tile * K0_TILE >= n) return;
    for (int k = tid; k < K0S_CHUNKS * p.n_rings; k += K0S_THREADS) chunk_cnt[k / p.n_rings][k % p.n_rings] = 0;
    int ring[K0S_U], rank[K0S_U]; float x[K0S_U], y[K0S_U], z[K0S_U], ori[K0S_U];
    #pragma unroll
    for (int u = 0; u < K0S_U; u++) {
        const int i = tile * K0_TILE + (warp * K0S_U + u) * 32 + lane;
        ring[u] = -1; x[u] = y[u] = z[u] = 0.f; ori[u] = 0.f;
        if (i < n) {
            const float *q = p.raw + (size_t)(o0 + i) * p.stride;
            k0_point(p, q, x[u], y[u], z[u]);
            ring[u] = ring_of[(size_t)b * p.N + i];
            if (ring[u] >= 0) ori[u] = ori_of[(size_t)b * p.N + i];
        }
    }
    __syncthreads();
    #pragma unroll
    for (int u = 0; u < K0S_U; u++) {
        const unsigned mask = __match_any_sync(0xffffffffu, ring[u]);
        rank[u] = __popc(mask & ((1u << lane) - 1u));
        if (ring[u] >= 0 && rank[u] == 0) chunk_cnt[warp * K0S_U + u][ring[u]] = __popc(mask);
    }
    __syncthreads();
    if (tid < p.n_rings) {
        int run = 0;
        #pragma unroll 8
        for (int c = 0; c < K0S_CHUNKS; c++) { int v = chunk_cnt[c][tid]; chunk_cnt[c][tid] = run; run += v; }
        run_start[tid + 1] = run;                                 // ring totals of the tile, scanned below
        run_dst[tid] = ring_start[b * (VLO_MAX_RINGS + 1) + tid] + tile_hist[((size_t)b * p.n_rings + tid) * p.tiles + tile];
    }
    __syncthreads();
    if (warp == 0) {
        // inclusive scan of the (up to 128) ring totals: 4 consecutive entries per lane + one shuffle scan
        int loc[4], sum = 0;
        #pragma unroll
        for (int e = 0; e < 4; e++) { const int r = lane * 4 + e; loc[e] = r < p.n_rings ? run_start[r + 1] : 0; sum += loc[e]; }
        int inc = sum;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
        int run = inc - sum;
        #pragma unroll
        for (int e = 0; e < 4; e++) { const int r = lane * 4 + e; run += loc[e]; if (r < p.n_rings) run_start[r + 1] = run; }
        if (lane == 0) run_start[0] = 0;
    }
    __syncthreads();
    const float startOri = ori_bounds[2 * b], endOri = ori_bounds[2 * b + 1];
    const int fh = first_half[b];
    #pragma unroll
    for (int u = 0; u < K0S_U; u++) {
        if (ring[u] < 0) continue;
        const int i = tile * K0_TILE + (warp * K0S_U + u) * 32 + lane;
        float o = ori[u];
        if (i <= fh) {
            if ((double)o < (double)startOri - VLO_PI_D / 2) o = (float)((double)o + 2 * VLO_PI_D);
            else if ((double)o > (double)startOri + VLO_PI_D * 3 / 2) o = (float)((double)o - 2 * VLO_PI_D);
        } else {
            o = (float)((double)o + 2 * VLO_PI_D);
            if ((double)o < (double)endOri - VLO_PI_D * 3 / 2) o = (float)((double)o + 2 * VLO_PI_D);
            else if ((double)o > (double)endOri + VLO_PI_D / 2) o = (float)((double)o - 2 * VLO_PI_D);
        }
        const float relTime = p.scan_period * (o - startOri) / (endOri - startOri);
        const int sp = run_start[ring[u]] + chunk_cnt[warp * K0S_U + u][ring[u]] + rank[u];
        sorted[sp] = make_float4(x[u], y[u], z[u], (float)ring[u] + relTime);
        sorted_src[sp] = i;
        sorted_ring[sp] = (int8_t)ring[u];
    }
    __syncthreads();
    const int n_out = run_start[p.n_rings];
    for (int k = tid; k < n_out; k += K0S_THREADS) {
        const int r = sorted_ring[k];
        const size_t pos = (size_t)b * p.N + run_dst[r] + (k - run_start[r]);
        cloud[pos] = sorted[k];
        src_index[pos] = sorted_src[k];
    }
}

int vlo_launch_organise(vlo_handle *h)
{
    ScanBatchDev &sb = h->sb;
    K0Params p;
    p.raw = sb.raw; p.raw_offset = sb.raw_offset; p.stride = sb.stride;
    p.xo = sb.xyz_off[0]; p.yo = sb.xyz_off[1]; p.zo = sb.xyz_off[2];
    p.n_rings = h->cfg.n_rings; p.lower_deg = h->cfg.lower_deg;
    p.factor = (float)(h->cfg.n_rings - 1) / (h->cfg.upper_deg - h->cfg.lower_deg);
    p.scan_period = h->cfg.scan_period; p.N = h->cfg.max_points; p.tiles = h->tiles_per_scan;
    {
        // same expression, in double, as the oracle's input_rotation_matrix
        const vlo_config &c = h->cfg;
        const double cy = cos((double)c.input_rotation[0]), sy = sin((double)c.input_rotation[0]);
        const double cp = cos((double)c.input_rotation[1]), sp = sin((double)c.input_rotation[1]);
        const double cr = cos((double)c.input_rotation[2]), sr = sin((double)c.input_rotation[2]);
        p.R[0] = (float)(cy * cp); p.R[1] = (float)(cy * sp * sr - sy * cr); p.R[2] = (float)(cy * sp * cr + sy * sr);
        p.R[3] = (float)(sy * cp); p.R[4] = (float)(sy * sp * sr + cy * cr); p.R[5] = (float)(sy * sp * cr - cy * sr);
        p.R[6] = (float)(-sp);     p.R[7] = (float)(cp * sr);                p.R[8] = (float)(cp * cr);
        p.rotate = c.rotate_input ? 1 : 0;
        p.ring_type = c.ring_field_type;
        const int ring_bytes = c.ring_field_type == 0 ? 4 : (c.ring_field_type == 1 ? 2 : 1);
        const int ring_off = c.ring_field_type == 0 ? c.ring_field * 4 : c.ring_field;
        p.ring_field = (c.ring_field >= 0 && c.ring_field_type >= 0 && c.ring_field_type <= 2 && ring_off + ring_bytes <= sb.stride * 4) ? c.ring_field : -1;
    }
    // One pass over the whole batch.  Sub-batches sized for the 126 MB L2 (pass 3 then re-reads the raw points and pass 1's
    // per-point results from L2 instead of DRAM) were measured and lost: 0.85 ms against 0.47 ms per 128 scans at 12 scans per
    // sub-batch, 0.61 ms at 24 -- these kernels need the whole batch in flight to cover their latency (VLO_K0_SUB keeps the knob).
    const int B = sb.scan_count;
    const int sub = h->k0_sub > 0 ? h->k0_sub : B;
    vlo_prof_begin(h, ST_ORGANISE);
    p.scan_first = sb.scan_first;
    k0_bounds<<<(B + 127) / 128, 128, 0, h->stream>>>(p, sb.ori_bounds, sb.first_half, B);
    h->launches += 1;
    for (int b0 = 0; b0 < B; b0 += sub) {
        const int nb = std::min(sub, B - b0);
        p.scan_first = sb.scan_first + b0;
        dim3 grid(h->tiles_per_scan, nb);
        k0_classify<<<grid, K0_TILE, 0, h->stream>>>(p, sb.ori_bounds, sb.first_half, sb.tile_hist, sb.ring_of, sb.ori_of);
        k0_scan<<<nb, 256, 0, h->stream>>>(p, sb.tile_hist, sb.ring_start, sb.counts);
        k0_scatter<<<grid, K0S_THREADS, 0, h->stream>>>(p, sb.ori_bounds, sb.first_half, sb.tile_hist, sb.ring_start,
                                                    sb.ring_of, sb.ori_of, sb.cloud, sb.src_index);
        h->launches += 3;
    }
    vlo_prof_end(h, ST_ORGANISE);
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}
