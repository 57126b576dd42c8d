// K3  scan-to-scan registration (LaserOdometry): association on the ring-segment box index of the previous sweep
// (segbox.cuh) + fused linearisation / deterministic reduction / on-device solve.
// Replaces BasicLaserOdometry::process of the `loam` nodelet (gtsam_fusion/launch/loam.launch:40-45;
// knobs loam_params.yaml:36-39); SURVEY.md Appendix A.4-A.7 is the algorithm, oracle/laser_odometry.c
// the frozen operation order (R1 three-level blocked summation, R2..R5).
//   k3_seg_build : boxes over the ring-major target clouds (32-point arcs, 32-arc groups), one CTA per (scan, cloud)
//   k3_assoc : one warp per feature point: transformToStart, exact 1-NN, ring-constrained partner
//              search with upstream's forward/backward tie order
//   k3_gn    : one CTA per scan pair runs up to 5 Gauss-Newton iterations between associations:
//              thread per correspondence -> 28 products -> shared-memory transposed R1 reduction ->
//              warp 0 solves (QR), iteration 0 also runs the single-warp Jacobi degeneracy test.
// The host enqueues [k3_assoc, k3_gn] x ceil(maxIter/5) back to back; convergence is a device flag,
// so control never returns to the host inside a registration.
#include "grid.cuh"
#include "segbox.cuh"
#ifndef K3_ONLINE_GRID
#define K3_ONLINE_GRID 1         // 1: a call of up to K3_THREAD_PAIRS pairs (the online tick) keeps round 1's voxel-hash warp kernel:
                                 // with the whole machine on ONE pair its latency is lower (HDL-64 tick p50 0.90 vs 1.50 ms)
#endif
#define K3_THREAD_PAIRS 4
#include "dense6.cuh"
#include "odom_lin.cuh"
#include <algorithm>

struct OdomParams {
    // features of the current sweep (queries) and the previous sweep (targets), per scan
    const float4 *sharp_pts, *flat_pts; int cap_sharp, cap_flat;
    const float4 *lsharp_pts; int cap_lsharp;            // dense ring-major, intensity = ring (+relTime)
    const float4 *lflat_pts; int N;                      // dense less-flat cloud [B][N]
    const int *lsharp_ring_start, *lflat_ring_start;     // [B][R+1]
    const int *counts;                                   // [B][8]
    int n_rings;
    // pairs
    const int *pair_last, *pair_cur; float *pair_T; int *pair_state; int *cidx, *sidx; vlo_result *result; int max_pairs;
    GridSet gc, gsf;                                     // corner / surf voxel-hash grids, grid index = scan index
    SegSet ss;                                           // ring-segment box index of every resident scan's target clouds (batches)
    int deskew; float inv_period; int fwd_quirk;
    int max_iter; float degen_thr, dT_abort, dR_abort, rot_thr, trans_thr;
};

__device__ __forceinline__ float4 lflat_point(const OdomParams &p, int scan, int dense)
{
    return p.lflat_pts[(size_t)scan * p.N + dense];
}

// ---- ring-segment box index of the target clouds (segbox.cuh): one CTA per (scan, cloud) --------------------------
#define SEGB_THREADS 512
struct SegBuildParams {
    const float4 *pts[2]; size_t stride[2]; const int *ring_start[2];      // dense ring-major clouds + [B][VLO_MAX_RINGS + 1]
    int n_rings, scan_first;
    SegSet ss;
};

__device__ __forceinline__ int seg_cell_of(const float4 lo, const float4 hi, const float *glo, const float *gsc)
{
    const float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
    const int ix = min(15, max(0, (int)((cx - glo[0]) * gsc[0]))), iy = min(3, max(0, (int)((cy - glo[1]) * gsc[1])));
    const int iz = min(15, max(0, (int)((cz - glo[2]) * gsc[2])));
    int m = 0;                                   // Morton order in the horizontal plane (x left, z forward), height fastest
    #pragma unroll
    for (int b = 0; b < 4; b++) m |= (((ix >> b) & 1) << (2 * b)) | (((iz >> b) & 1) << (2 * b + 1));
    return (m << 2) | iy;
}

__device__ __forceinline__ float seg_warp_min(float v) {
    #pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = fminf(v, __shfl_down_sync(0xffffffffu, v, d));
    return v;
}
__device__ __forceinline__ float seg_warp_max(float v) {
    #pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, d));
    return v;
}

// Where upstream's partner loops break (oracle/laser_odometry.c partners_corner / partners_surf): started at target
// point j whose scan id is e = int(w_j), the forward loop stops at the first k > j with int(w_k) > e + 2.5 and the backward loop
// at the first k < j with int(w_k) < e - 2.5; out[j] = (that backward k or -1, that forward k or n, and -- with the box index -- the arcs
// [f0, f1) that hold the points in between).  Scan ids follow the
// ring-major order except where relTime is negative (segbox.cuh), so the answer is nearly always the start of ring e + 3 / the
// end of ring e - 3: rings whose min / max scan id (s_emin / s_emax, filled by the caller) cannot stop the loop are skipped
// whole, the stopping ring is walked.  No assumption on the scan ids beyond that; CTA-wide, no barrier inside.
// the two break positions for loops started at point j of ring slot `lo` whose scan id is e
__device__ __forceinline__ void partner_breaks(const float4 *pts, const int *rs, int R, const int *s_emin, const int *s_emax,
                                               int j_fwd, int j_bwd, int lo, int e, int &B, int &F)
{
    F = rs[R]; B = -1;
    {
        const int t = e + 3; int k = j_fwd + 1; bool found = false;
        for (int r = lo; r < R && !found; r++) {
            if (s_emax[r] >= t) {
                const int end = rs[r + 1];
                for (k = max(k, rs[r]); k < end; k++) if ((int)pts[k].w >= t) { F = k; found = true; break; }
            }
        }
    }
    {
        const int t = e - 3; int k = j_bwd - 1; bool found = false;
        for (int r = lo; r >= 0 && !found; r--) {
            if (s_emin[r] <= t) {
                const int beg = rs[r];
                for (k = min(k, rs[r + 1] - 1); k >= beg; k--) if ((int)pts[k].w <= t) { B = k; found = true; break; }
            }
        }
    }
}
// the arcs [f0, f1) of the box index that hold the points between the two break positions
__device__ __forceinline__ int4 partner_range_record(const int *rs, int R, const int *seg_ring, int B, int F)
{
    int f0 = 0, f1 = 0;
    if (seg_ring && F - B > 1) {
        int a = 0, b = R;
        while (b - a > 1) { const int mid = (a + b) >> 1; if (rs[mid] <= B + 1) a = mid; else b = mid; }
        f0 = seg_ring[a] + ((B + 1 - rs[a]) >> SEG_SHIFT);
        a = 0; b = R;
        while (b - a > 1) { const int mid = (a + b) >> 1; if (rs[mid] <= F - 1) a = mid; else b = mid; }
        f1 = seg_ring[a] + ((F - 1 - rs[a]) >> SEG_SHIFT) + 1;
    }
    return make_int4(B, F, f0, f1);
}
// A warp per ring.  All points of a ring whose scan ids agree (min == max: every ring of a sweep without negative relTime) share
// one record -- no point of the ring itself can stop a loop started in it -- computed once; otherwise point by point.
__device__ __forceinline__ void partner_ranges_cta(const float4 *pts, const int *rs, int R, int4 *out, const int *s_emin, const int *s_emax,
                                                   const int *seg_ring, int tid, int n_threads)
{
    const int lane = tid & 31;
    for (int r = tid >> 5; r < R; r += n_threads >> 5) {
        const int j0 = rs[r], j1 = rs[r + 1];
        if (j1 <= j0) continue;
        if (s_emin[r] == s_emax[r]) {
            int4 rec = make_int4(0, 0, 0, 0);
            if (lane == 0) {
                int B, F;
                partner_breaks(pts, rs, R, s_emin, s_emax, j1 - 1, j0, r, s_emin[r], B, F);
                rec = partner_range_record(rs, R, seg_ring, B, F);
            }
            rec.x = __shfl_sync(0xffffffffu, rec.x, 0); rec.y = __shfl_sync(0xffffffffu, rec.y, 0);
            rec.z = __shfl_sync(0xffffffffu, rec.z, 0); rec.w = __shfl_sync(0xffffffffu, rec.w, 0);
            for (int j = j0 + lane; j < j1; j += 32) out[j] = rec;
        } else {
            for (int j = j0 + lane; j < j1; j += 32) {
                int B, F;
                partner_breaks(pts, rs, R, s_emin, s_emax, j, j, r, (int)pts[j].w, B, F);
                out[j] = partner_range_record(rs, R, seg_ring, B, F);
            }
        }
    }
}

// min / max scan id of every ring slot into shared memory (the caller barriers before and after)
__device__ __forceinline__ void ring_scan_extrema_cta(const float4 *pts, const int *rs, int R, int *s_emin, int *s_emax, int tid, int n_threads)
{
    const int lane = tid & 31;
    for (int r = tid >> 5; r < R; r += n_threads >> 5) {        // a warp per ring
        int mn = 0x7fffffff, mx = (int)0x80000000;
        for (int k = rs[r] + lane; k < rs[r + 1]; k += 32) { const int e = (int)pts[k].w; mn = min(mn, e); mx = max(mx, e); }
        mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) { s_emin[r] = mn; s_emax[r] = mx; }
    }
}

#if K3_ONLINE_GRID
// the partner-loop ranges alone (the online tick searches voxel-hash grids, not the box index): one CTA per (scan, cloud)
__global__ void __launch_bounds__(SEGB_THREADS) k3_partner_ranges(SegBuildParams p)
{
    __shared__ int s_emin[VLO_MAX_RINGS], s_emax[VLO_MAX_RINGS];
    const int b = p.scan_first + blockIdx.x, w = blockIdx.y, tid = threadIdx.x;
    const float4 *pts = p.pts[w] + (size_t)b * p.stride[w];
    const int *rs = p.ring_start[w] + b * (VLO_MAX_RINGS + 1);
    ring_scan_extrema_cta(pts, rs, p.n_rings, s_emin, s_emax, tid, SEGB_THREADS);
    __syncthreads();
    partner_ranges_cta(pts, rs, p.n_rings, p.ss.prange[w] + (size_t)b * p.stride[w], s_emin, s_emax, nullptr, tid, SEGB_THREADS);
}
#endif

__global__ void __launch_bounds__(SEGB_THREADS) k3_seg_build(SegBuildParams p)
{
    __shared__ int s_emin[VLO_MAX_RINGS], s_emax[VLO_MAX_RINGS];
    __shared__ int s_seg_ring[VLO_MAX_RINGS + 1];
    __shared__ int s_hist[SEG_CELLS];
    __shared__ int s_wsum[SEGB_THREADS / 32];
    __shared__ float s_red[6][SEGB_THREADS / 32];
    __shared__ float s_glo[3], s_gsc[3];
    const int b = p.scan_first + blockIdx.x, w = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float4 *pts = p.pts[w] + (size_t)b * p.stride[w];
    const int *rs = p.ring_start[w] + b * (VLO_MAX_RINGS + 1);
    float4 *fbox = p.ss.fbox[w] + (size_t)b * p.ss.max_seg[w] * 2;
    float4 *cbox = p.ss.cbox[w] + (size_t)b * p.ss.max_coarse[w] * 2;
    int *perm = p.ss.perm[w] + (size_t)b * p.ss.max_seg[w];
    float4 *mbox = p.ss.mbox[w] + (size_t)b * p.ss.max_seg[w] * 2;
    int *seg_ring = p.ss.seg_ring[w] + b * (VLO_MAX_RINGS + 1);
    const float INF = __int_as_float(0x7f800000);
    // 1. segments per ring (ceil(n_r / SEG_PTS)), exclusive prefix
    if (warp == 0) {
        int carry = 0;
        for (int r0 = 0; r0 < p.n_rings; r0 += 32) {
            const int r = r0 + lane;
            const int v = r < p.n_rings ? (rs[r + 1] - rs[r] + SEG_PTS - 1) / SEG_PTS : 0;
            int inc = v;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
            if (r < p.n_rings) s_seg_ring[r] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_seg_ring[p.n_rings] = carry;
    }
    for (int k = tid; k < SEG_CELLS; k += SEGB_THREADS) s_hist[k] = 0;
    __syncthreads();
    const int nseg = min(s_seg_ring[p.n_rings], p.ss.max_seg[w]);
    for (int r = tid; r <= VLO_MAX_RINGS; r += SEGB_THREADS) seg_ring[r] = min(s_seg_ring[min(r, p.n_rings)], nseg);
    if (tid == 0) p.ss.nseg[w][b] = nseg;
    // 2. fine boxes: a lane group of SEG_PTS lanes per segment
    {
        const int g = lane >> SEG_SHIFT, l = lane & (SEG_PTS - 1);
        for (int fb = warp * SEG_GROUPS; fb < nseg; fb += (SEGB_THREADS / 32) * SEG_GROUPS) {       // warp-uniform
            const int f = fb + g;
            int lo = 0, hi = p.n_rings, s0 = 0, cnt = 0;
            if (f < nseg) {
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_seg_ring[mid] <= f) lo = mid; else hi = mid; }   // ring of segment f
                s0 = rs[lo] + SEG_PTS * (f - s_seg_ring[lo]); cnt = min(SEG_PTS, rs[lo + 1] - s0);
            }
            float x0 = INF, y0 = INF, z0 = INF, x1 = -INF, y1 = -INF, z1 = -INF;
            int e0 = 0x7fffffff, e1 = (int)0x80000000;           // min / max scan id (int(w), segbox.cuh) of the arc
            if (l < cnt) { const float4 q = pts[s0 + l]; x0 = x1 = q.x; y0 = y1 = q.y; z0 = z1 = q.z; e0 = e1 = (int)q.w; }
            #pragma unroll
            for (int d = SEG_PTS / 2; d >= 1; d >>= 1) {       // the group's lane 0 ends up with the group's extrema
                x0 = fminf(x0, __shfl_down_sync(0xffffffffu, x0, d)); y0 = fminf(y0, __shfl_down_sync(0xffffffffu, y0, d));
                z0 = fminf(z0, __shfl_down_sync(0xffffffffu, z0, d)); x1 = fmaxf(x1, __shfl_down_sync(0xffffffffu, x1, d));
                y1 = fmaxf(y1, __shfl_down_sync(0xffffffffu, y1, d)); z1 = fmaxf(z1, __shfl_down_sync(0xffffffffu, z1, d));
                e0 = min(e0, __shfl_down_sync(0xffffffffu, e0, d)); e1 = max(e1, __shfl_down_sync(0xffffffffu, e1, d));
            }
            if (l == 0 && f < nseg) {
                fbox[2 * f] = make_float4(x0, y0, z0, __int_as_float(s0));
                fbox[2 * f + 1] = make_float4(x1, y1, z1, __int_as_float(SEG_META(e0, e1, cnt)));
            }
        }
    }
    __syncthreads();
    // bounding box of the box centres
    {
        float blo[3] = { INF, INF, INF }, bhi[3] = { -INF, -INF, -INF };
        for (int f = tid; f < nseg; f += SEGB_THREADS) {
            const float4 lo = fbox[2 * f], hi = fbox[2 * f + 1];
            const float c[3] = { 0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z) };
            #pragma unroll
            for (int a = 0; a < 3; a++) { blo[a] = fminf(blo[a], c[a]); bhi[a] = fmaxf(bhi[a], c[a]); }
        }
        #pragma unroll
        for (int a = 0; a < 3; a++) { blo[a] = seg_warp_min(blo[a]); bhi[a] = seg_warp_max(bhi[a]); }
        if (lane == 0) {
            #pragma unroll
            for (int a = 0; a < 3; a++) { s_red[a][warp] = blo[a]; s_red[3 + a][warp] = bhi[a]; }
        }
    }
    __syncthreads();
    if (tid < 3) {
        float lo = INF, hi = -INF;
        for (int k = 0; k < SEGB_THREADS / 32; k++) { lo = fminf(lo, s_red[tid][k]); hi = fmaxf(hi, s_red[3 + tid][k]); }
        const float ext = hi - lo;
        s_glo[tid] = lo;
        s_gsc[tid] = (ext > 1e-6f) ? (tid == 1 ? 4.0f : 16.0f) / ext : 0.0f;
    }
    __syncthreads();
    // 3. counting sort of the segments by the cell of their box centre -> perm
    for (int f = tid; f < nseg; f += SEGB_THREADS) atomicAdd(&s_hist[seg_cell_of(fbox[2 * f], fbox[2 * f + 1], s_glo, s_gsc)], 1);
    __syncthreads();
    {
        const int k0 = tid * (SEG_CELLS / SEGB_THREADS);         // consecutive cells per thread
        int loc[SEG_CELLS / SEGB_THREADS], sum = 0;
        #pragma unroll
        for (int e = 0; e < SEG_CELLS / SEGB_THREADS; e++) { loc[e] = s_hist[k0 + e]; sum += loc[e]; }
        int inc = sum;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        int off = 0;
        for (int k = 0; k < warp; k++) off += s_wsum[k];
        int run = off + inc - sum;
        #pragma unroll
        for (int e = 0; e < SEG_CELLS / SEGB_THREADS; e++) { s_hist[k0 + e] = run; run += loc[e]; }
    }
    __syncthreads();
    for (int f = tid; f < nseg; f += SEGB_THREADS)
        perm[atomicAdd(&s_hist[seg_cell_of(fbox[2 * f], fbox[2 * f + 1], s_glo, s_gsc)], 1)] = f;
    __syncthreads();
    // 4. coarse boxes: group c = perm[32 c .. 32 c + 32)
    const int ncoarse = (nseg + 31) >> 5;
    for (int c = warp; c < ncoarse; c += SEGB_THREADS / 32) {
        const int nmem = min(32, nseg - 32 * c);
        float x0 = INF, y0 = INF, z0 = INF, x1 = -INF, y1 = -INF, z1 = -INF;
        if (lane < nmem) {
            const int f = perm[32 * c + lane];
            const float4 lo = fbox[2 * f], hi = fbox[2 * f + 1];
            mbox[2 * (32 * c + lane)] = lo; mbox[2 * (32 * c + lane) + 1] = hi;       // the group's members, side by side
            x0 = lo.x; y0 = lo.y; z0 = lo.z; x1 = hi.x; y1 = hi.y; z1 = hi.z;
        }
        x0 = seg_warp_min(x0); y0 = seg_warp_min(y0); z0 = seg_warp_min(z0);
        x1 = seg_warp_max(x1); y1 = seg_warp_max(y1); z1 = seg_warp_max(z1);
        if (lane == 0) { cbox[2 * c] = make_float4(x0, y0, z0, __int_as_float(nmem)); cbox[2 * c + 1] = make_float4(x1, y1, z1, 0.0f); }
    }
    // 5. where the partner loops started at each point break
    ring_scan_extrema_cta(pts, rs, p.n_rings, s_emin, s_emax, tid, SEGB_THREADS);
    __syncthreads();
    partner_ranges_cta(pts, rs, p.n_rings, p.ss.prange[w] + (size_t)b * p.stride[w], s_emin, s_emax, s_seg_ring, tid, SEGB_THREADS);
}

__device__ __forceinline__ SegCloud seg_cloud_of(const OdomParams &p, int w, int scan)
{
    SegCloud c;
    c.pts = (w == 0 ? p.lsharp_pts + (size_t)scan * p.cap_lsharp : p.lflat_pts + (size_t)scan * p.N);
    c.fbox = p.ss.fbox[w] + (size_t)scan * p.ss.max_seg[w] * 2;
    c.cbox = p.ss.cbox[w] + (size_t)scan * p.ss.max_coarse[w] * 2;
    c.mbox = p.ss.mbox[w] + (size_t)scan * p.ss.max_seg[w] * 2;
    c.seg_ring = p.ss.seg_ring[w] + scan * (VLO_MAX_RINGS + 1);
    c.ring_start = (w == 0 ? p.lsharp_ring_start : p.lflat_ring_start) + scan * (VLO_MAX_RINGS + 1);
    c.prange = p.ss.prange[w] + (size_t)scan * (w == 0 ? p.cap_lsharp : p.N);
    c.nseg = p.ss.nseg[w][scan]; c.ncoarse = (c.nseg + 31) >> 5; c.n_rings = p.n_rings;
    return c;
}

#if K3_ONLINE_GRID
// Online tick (a few pairs): one warp per feature point on the voxel-hash grids (grid.cuh grid_search: shells of cells
// resolved cooperatively), the whole machine on one pair.
__global__ void __launch_bounds__(256) k3_assoc_warp(OdomParams p, int n_pairs)
{
    __shared__ int scratch[8][GRID_SCRATCH_INTS];
    const int pair = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    if (p.pair_state[pair * 4 + 0]) return;                          // converged
    const int last = p.pair_last[pair], cur = p.pair_cur[pair];
    const int n_sharp = p.counts[cur * 8 + 1], n_flat = p.counts[cur * 8 + 3];
    const int n_lc = p.counts[last * 8 + 2], n_ls = p.counts[last * 8 + 4];
    if (!(n_lc > 10 && n_ls > 100)) return;
    float T[6];
    #pragma unroll
    for (int a = 0; a < 6; a++) T[a] = p.pair_T[pair * 6 + a];
    // persistent warps stride over the compact query index space [sharp..., flat...]
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_sharp + n_flat; w += n_warps) {
        if (w < n_sharp) {
            float4 q = vlo_to_start(T, p.sharp_pts[(size_t)cur * p.cap_sharp + w], p.deskew, p.inv_period);
            TopK<1> nn;
            grid_search<1>(p.gc, last, q.x, q.y, q.z, 25.0f, FilterAll(), nn, lane, scratch[warp]);
            int i1 = -1, i2 = -1;
            if (nn.tag[0] != GRID_NOTAG) {
                i1 = (int)(nn.tag[0] & 0xFFFFFFu);
                const int4 rg = p.ss.prange[0][(size_t)last * p.cap_lsharp + i1];
                FilterPartnerRange f; f.ind = i1; f.scan = (int)(nn.tag[0] >> 24); f.lo = rg.x; f.want = 3;
                f.hi = min(rg.y, p.fwd_quirk ? min(n_sharp, n_lc) : n_lc);
                TopK<1> pr;
                grid_search<1>(p.gc, last, q.x, q.y, q.z, 25.0f, f, pr, lane, scratch[warp]);
                if (pr.tag[0] != GRID_NOTAG) i2 = (int)(pr.tag[0] & 0xFFFFFFu);
            }
            if (lane == 0) {
                int *o = p.cidx + ((size_t)pair * p.cap_sharp + w) * 2;
                o[0] = i1; o[1] = i2;
            }
        } else {
            int f_i = w - n_sharp;
            float4 q = vlo_to_start(T, p.flat_pts[(size_t)cur * p.cap_flat + f_i], p.deskew, p.inv_period);
            TopK<1> nn;
            grid_search<1>(p.gsf, last, q.x, q.y, q.z, 25.0f, FilterAll(), nn, lane, scratch[warp]);
            int i1 = -1, i2 = -1, i3 = -1;
            if (nn.tag[0] != GRID_NOTAG) {
                i1 = (int)(nn.tag[0] & 0xFFFFFFu);
                const int4 rg = p.ss.prange[1][(size_t)last * p.N + i1];
                FilterPartnerRange f2; f2.ind = i1; f2.scan = (int)(nn.tag[0] >> 24); f2.lo = rg.x; f2.want = 2;
                f2.hi = min(rg.y, p.fwd_quirk ? min(n_flat, n_ls) : n_ls);
                FilterPartnerRange f3 = f2; f3.want = 3;
                TopK<1> p2, p3;
                grid_search<1>(p.gsf, last, q.x, q.y, q.z, 25.0f, f2, p2, lane, scratch[warp]);
                grid_search<1>(p.gsf, last, q.x, q.y, q.z, 25.0f, f3, p3, lane, scratch[warp]);
                if (p2.tag[0] != GRID_NOTAG) i2 = (int)(p2.tag[0] & 0xFFFFFFu);
                if (p3.tag[0] != GRID_NOTAG) i3 = (int)(p3.tag[0] & 0xFFFFFFu);
            }
            if (lane == 0) {
                int *o = p.sidx + ((size_t)pair * p.cap_flat + f_i) * 3;
                o[0] = i1; o[1] = i2; o[2] = i3;
            }
        }
    }
}

#endif

// Association on the ring-segment box index: one WARP per feature point of the current sweep -- transformToStart, exact
// nearest neighbour in the previous sweep's cloud (d2 < 25), ring-constrained partner search(es) with upstream's forward /
// backward tie order.  From the second association round on, the previous round's answers (same pair, pose a few Gauss-Newton
// steps older) seed the searches: they are candidates like any other, and their distances prune nearly every box at once.
// Measured against round 1's voxel-hash kernels on the whole-bag workload (127 HDL-64 pairs, 5 rounds, profiles/r02*):
// 4.9 ms vs 11.2 / 22.8 ms (1.0 / 0.7 m cells) -- the hash version fell back to walking cell shells whenever a partner was
// more than one cell edge away -- with bit-identical results.
#define K3A_THREADS 128
#ifndef K3A_WQ
#define K3A_WQ 4                 // queries per ticket: a WARP draws (pair, 4 consecutive queries) tickets (3.35 ms per 127-pair step; 8: 3.46, 16: 3.80, 1: 3.57)
#endif
#ifndef K3A_MINB
#define K3A_MINB 16          // 32 registers (148 B of spills): all 64 warp slots of an SM.  255-pair step: 6.89 ms at 8 CTAs per SM (64 registers),
                             // 6.51 at 10, 6.19 at 12 (40 registers, no spill), 6.06 at 16; left to itself ptxas takes 128 registers: 1.6x slower
#endif
// Persistent warps, no CTA-level step: a ticket is (pair, K3A_WQ queries); the warp's first lanes transform the ticket's queries
// (transformToStart, one query per lane), then the warp answers them one after the other.  Measured on the 127-pair step: a
// (share, pair) grid with the share's queries dealt to the CTA's warps in turn and a barrier per share took 4.49 ms at 128
// queries per share, 3.71 at 64, 3.40 at 16 -- the barrier waited for the warp with the expensive queries; warp tickets remove it.
__global__ void __launch_bounds__(K3A_THREADS, K3A_MINB) k3_assoc(OdomParams p, int n_pairs, int round, int shares_per_pair)
{
    const int lane = threadIdx.x & 31;
    int *tickets = p.pair_state + (size_t)p.max_pairs * 4;           // one work counter per association round (zeroed by k3_init_pairs)
    while (true) {
        int ticket = 0;
        if (lane == 0) ticket = atomicAdd(&tickets[round], 1);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket >= n_pairs * shares_per_pair) break;
        const int pair = ticket / shares_per_pair, base = (ticket - pair * shares_per_pair) * K3A_WQ;
        if (p.pair_state[pair * 4 + 0]) continue;                    // converged
        const int last = p.pair_last[pair], cur = p.pair_cur[pair];
        const int n_sharp = p.counts[cur * 8 + 1], n_flat = p.counts[cur * 8 + 3];
        const int n_lc = p.counts[last * 8 + 2], n_ls = p.counts[last * 8 + 4];
        if (!(n_lc > 10 && n_ls > 100)) continue;
        const int total = n_sharp + n_flat;
        if (base >= total) continue;
        const int nq = min(K3A_WQ, total - base);
        float4 myq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < nq) {
            float T[6];
            #pragma unroll
            for (int a = 0; a < 6; a++) T[a] = p.pair_T[pair * 6 + a];
            const int wq = base + lane;
            myq = vlo_to_start(T, wq < n_sharp ? p.sharp_pts[(size_t)cur * p.cap_sharp + wq] : p.flat_pts[(size_t)cur * p.cap_flat + (wq - n_sharp)],
                               p.deskew, p.inv_period);
        }
        for (int k = 0; k < nq; k++) {
            const int wq = base + k;
            float4 q;
            q.x = __shfl_sync(0xffffffffu, myq.x, k); q.y = __shfl_sync(0xffffffffu, myq.y, k); q.z = __shfl_sync(0xffffffffu, myq.z, k);
            const bool sharp = wq < n_sharp;
            const SegCloud c = seg_cloud_of(p, sharp ? 0 : 1, last);
            int *o = sharp ? p.cidx + ((size_t)pair * p.cap_sharp + wq) * 2 : p.sidx + ((size_t)pair * p.cap_flat + (wq - n_sharp)) * 3;
            int s1 = -1, s2 = -1, s3 = -1;
            if (round > 0) { s1 = o[0]; s2 = o[1]; if (!sharp) s3 = o[2]; }
            const int n_tgt = sharp ? n_lc : n_ls;
            SegFilter f; f.mode = 0; f.ind = -1; f.scan = 0; f.lo = 0; f.hi = 0; f.want = 0; f.f0 = f.f1 = 0;
            int scan = 0;
            const int i1 = seg_search(c, q.x, q.y, q.z, 25.0f, f, (s1 >= 0 && s1 < n_tgt) ? s1 : -1, lane, &scan);
            int i2 = -1, i3 = -1;
            if (i1 >= 0) {
                const int4 pr = c.prange[i1];                                // where the loops started at i1 break, the arcs in between
                f.mode = 1; f.ind = i1; f.scan = scan; f.lo = pr.x; f.want = 3; f.f0 = pr.z; f.f1 = pr.w;
                f.hi = min(pr.y, sharp ? (p.fwd_quirk ? min(n_sharp, n_lc) : n_lc) : (p.fwd_quirk ? min(n_flat, n_ls) : n_ls));
                if (sharp) {
                    i2 = seg_search(c, q.x, q.y, q.z, 25.0f, f, (s2 >= 0 && s2 < n_tgt) ? s2 : -1, lane, nullptr);
                } else {
                    seg_search_partners(c, q.x, q.y, q.z, 25.0f, f, (s2 >= 0 && s2 < n_tgt) ? s2 : -1,          // same-scan and other-scan partner
                                        (s3 >= 0 && s3 < n_tgt) ? s3 : -1, lane, i2, i3);
                }
            }
            if (lane == 0) { o[0] = i1; o[1] = i2; if (!sharp) o[2] = i3; }
        }
    }
}

#ifndef GN_THREADS
#define GN_THREADS 512     // one CTA per pair: the pair's 2304 correspondences in 5 chunks (8 warps / 9 chunks before: 26 us per iteration on one SM)
#endif
#define GN_GROUPS (GN_THREADS / 32)
#define TSTRIDE 29     // padded row of the transposed term buffer

// R1 block reduction of one chunk of GN_THREADS queries held in smem terms[GN_THREADS][TSTRIDE] (the comments count for 256):
// 224 threads sum 32 consecutive queries each (level 1), then 28 threads fold the 8 level-1 sums into
// the running level-2 accumulator; level 2 closes into level 3 every 1024 queries.
__device__ __forceinline__ void r1_chunk(float *terms, float *l1buf, float *l2acc, float *l3acc, int chunk, int n_chunks, int q_total, int tid)
{
    __syncthreads();
    for (int task = tid; task < GN_GROUPS * VLO_NTERM; task += GN_THREADS) {
        int g = task / VLO_NTERM, e = task % VLO_NTERM;
        float l1 = 0.0f;
        const float *src = terms + (size_t)(g * 32) * TSTRIDE + e;
        #pragma unroll 8
        for (int k = 0; k < 32; k++) l1 = l1 + src[(size_t)k * TSTRIDE];
        l1buf[g * VLO_NTERM + e] = l1;
    }
    __syncthreads();
    if (tid < VLO_NTERM) {
        float l2 = l2acc[tid];
        for (int g = 0; g < GN_GROUPS; g++) {
            const int q0 = chunk * GN_THREADS + g * 32;          // level 1 block = 32 queries, level 2 closes every 1024 queries
            if (q0 < q_total) l2 = l2 + l1buf[g * VLO_NTERM + tid];
            if (q0 < q_total && ((q0 & 1023) == 992 || q0 + 32 >= q_total)) { l3acc[tid] = l3acc[tid] + l2; l2 = 0.0f; }
        }
        l2acc[tid] = l2;
    }
}

#ifndef GN_MINB
#define GN_MINB 2          // 64 registers, two CTAs per SM: the 255 pairs of a 256-frame batch run in one wave (k3_gn 1.17 -> 0.82 ms; 127 pairs: 0.60 -> 0.63)
#endif
__global__ void __launch_bounds__(GN_THREADS, GN_MINB) k3_gn(OdomParams p, int iter_base, int n_iters)
{
#ifdef VLO_HOST_EMULATION
    static float terms[GN_THREADS * TSTRIDE];
#else
    extern __shared__ float terms[];             // [GN_THREADS][TSTRIDE]: 58 KB at 512 threads (dynamic, opted in per handle)
#endif
    __shared__ float l1buf[GN_GROUPS * VLO_NTERM];
    __shared__ float l2acc[VLO_NTERM], l3acc[VLO_NTERM];
    __shared__ float trig[6];
    __shared__ GnScratch S;
    __shared__ int s_ne, s_np;
    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int *state = p.pair_state + pair * 4;      // converged, iterations, is_degenerate, status
    if (state[0]) return;
    const int last = p.pair_last[pair], cur = p.pair_cur[pair];
    const int n_sharp = p.counts[cur * 8 + 1], n_flat = p.counts[cur * 8 + 3];
    const int n_lc = p.counts[last * 8 + 2], n_ls = p.counts[last * 8 + 4];
    vlo_result *res = p.result + pair;
    if (!(n_lc > 10 && n_ls > 100)) {           // upstream: not enough points in the last clouds -> no optimisation
        if (tid == 0 && iter_base == 0) { state[1] = 0; state[3] = VLO_SOFT_TOO_FEW_CORR; res->status = VLO_SOFT_TOO_FEW_CORR; res->iterations = 0; }
        return;
    }
    if (tid < 6) S.T[tid] = p.pair_T[pair * 6 + tid];
    if (tid == 0) {
        S.is_degenerate = state[2]; S.converged = 0; S.status = state[3]; S.iterations = state[1];
        S.n_edge = res->n_corr_edge; S.n_plane = res->n_corr_plane;
    }
    if (tid < 36 && iter_base > 0) S.P[tid] = res->P[tid];
    if (tid < VLO_NTERM) S.total[tid] = 0.0f;
    __syncthreads();
    const int q_total = n_sharp + n_flat;
    const int n_chunks = (q_total + GN_THREADS - 1) / GN_THREADS;
    bool have_total = false, did_eig = false;
    for (int it = iter_base; it < iter_base + n_iters && it < p.max_iter; it++) {
        if (tid == 0) {
            vlo_sincosf(S.T[0], trig[0], trig[1]);
            vlo_sincosf(S.T[1], trig[2], trig[3]);
            vlo_sincosf(S.T[2], trig[4], trig[5]);
            s_ne = 0; s_np = 0;
        }
        if (tid < VLO_NTERM) { l2acc[tid] = 0.0f; l3acc[tid] = 0.0f; }
        __syncthreads();
        float T[6];
        #pragma unroll
        for (int a = 0; a < 6; a++) T[a] = S.T[a];
        int my_edge = 0, my_plane = 0;
        for (int chunk = 0; chunk < n_chunks; chunk++) {
            int i = chunk * GN_THREADS + tid;
            float t[VLO_NTERM];
            #pragma unroll
            for (int e = 0; e < VLO_NTERM; e++) t[e] = 0.0f;
            if (i < q_total) {
                float coeff[4]; bool keep = false; float4 ori;
                if (i < n_sharp) {
                    const int *ci = p.cidx + ((size_t)pair * p.cap_sharp + i) * 2;
                    int i1 = ci[0], i2 = ci[1];
                    ori = p.sharp_pts[(size_t)cur * p.cap_sharp + i];
                    if (i2 >= 0) {
                        float4 sel = vlo_to_start(T, ori, p.deskew, p.inv_period);
                        float4 a = p.lsharp_pts[(size_t)last * p.cap_lsharp + i1];
                        float4 b = p.lsharp_pts[(size_t)last * p.cap_lsharp + i2];
                        keep = edge_coeff(sel, a, b, it, coeff);
                        if (keep) my_edge++;
                    }
                } else {
                    int f_i = i - n_sharp;
                    const int *si = p.sidx + ((size_t)pair * p.cap_flat + f_i) * 3;
                    int i1 = si[0], i2 = si[1], i3 = si[2];
                    ori = p.flat_pts[(size_t)cur * p.cap_flat + f_i];
                    if (i2 >= 0 && i3 >= 0) {
                        float4 sel = vlo_to_start(T, ori, p.deskew, p.inv_period);
                        float4 t1 = lflat_point(p, last, i1), t2 = lflat_point(p, last, i2), t3 = lflat_point(p, last, i3);
                        keep = plane_coeff(sel, t1, t2, t3, it, coeff);
                        if (keep) my_plane++;
                    }
                }
                if (keep) {
                    float row[6], bval;
                    odom_jacobian_row(T, trig, ori.x, ori.y, ori.z, coeff, row, bval);
                    int e = 0;
                    #pragma unroll
                    for (int a = 0; a < 6; a++)
                        #pragma unroll
                        for (int b = a; b < 6; b++) t[e++] = row[a] * row[b];
                    #pragma unroll
                    for (int a = 0; a < 6; a++) t[e++] = row[a] * bval;
                    t[e] = coeff[3] * coeff[3];
                }
            }
            __syncthreads();     // previous chunk's level-1 reads are done
            #pragma unroll
            for (int e = 0; e < VLO_NTERM; e++) terms[tid * TSTRIDE + e] = t[e];
            r1_chunk(terms, l1buf, l2acc, l3acc, chunk, n_chunks, q_total, tid);
        }
        // correspondence counts
        unsigned be = __reduce_add_sync(0xffffffffu, my_edge), bp = __reduce_add_sync(0xffffffffu, my_plane);
        if (lane == 0) { atomicAdd(&s_ne, (int)be); atomicAdd(&s_np, (int)bp); }
        __syncthreads();
        if (tid == 0) S.iterations = it + 1;
        if (s_ne + s_np < 10) { __syncthreads(); continue; }     // upstream: `continue` without update
        if (tid < VLO_NTERM) S.total[tid] = l3acc[tid];
        if (tid == 0) { S.n_edge = s_ne; S.n_plane = s_np; S.status = VLO_OK; }
        have_total = true;
        if (it == 0) did_eig = true;
        __syncthreads();
        if (warp == 0) vlo_gn_update_warp(S, it, p.degen_thr, p.dT_abort, p.dR_abort, lane);
        __syncthreads();
        if (S.converged) break;
    }
    __syncthreads();
    if (tid < 6) { p.pair_T[pair * 6 + tid] = S.T[tid]; res->transform[tid] = S.T[tid]; }
    if (did_eig) {
        if (tid < 36) res->P[tid] = S.P[tid];
        if (tid < 6) res->eig[tid] = S.eval[tid];
    }
    if (tid == 0) {
        state[0] = S.converged; state[1] = S.iterations; state[2] = S.is_degenerate; state[3] = S.status;
        res->iterations = S.iterations; res->is_degenerate = S.is_degenerate; res->status = S.status;
        if (have_total) {
            res->n_corr_edge = S.n_edge; res->n_corr_plane = S.n_plane;
            vlo_finish_result(S, p.rot_thr, p.trans_thr, res);
        }
    }
}

// reset per-pair state and seed the transform
__global__ void k3_init_pairs(OdomParams p, const float *seeds, int n_pairs)
{
    int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair < 64) p.pair_state[(size_t)p.max_pairs * 4 + pair] = 0;            // the association rounds' work tickets
    if (pair >= n_pairs) return;
    for (int a = 0; a < 6; a++) p.pair_T[pair * 6 + a] = seeds ? seeds[pair * 6 + a] : 0.0f;
    int *st = p.pair_state + pair * 4;
    st[0] = 0; st[1] = 0; st[2] = 0; st[3] = VLO_SOFT_TOO_FEW_CORR;
    vlo_result *r = p.result + pair;
    for (int a = 0; a < 6; a++) { r->transform[a] = p.pair_T[pair * 6 + a]; r->eig[a] = 0.0f; }
    for (int a = 0; a < 36; a++) { r->hessian[a] = 0.0f; r->P[a] = (a % 7 == 0) ? 1.0f : 0.0f; r->cov[a] = 0.0; }
    r->is_degenerate = 0; r->iterations = 0; r->n_corr_edge = 0; r->n_corr_plane = 0;
    r->logdet_rot = 0.f; r->logdet_trans = 0.f; r->pass_dopt = 0; r->status = VLO_SOFT_TOO_FEW_CORR;
}

// transformToEnd applied in place to the previous sweep's target clouds (online mode)
__global__ void __launch_bounds__(256) k3_to_end(OdomParams p, const int *scans, const float *Ts, int n, int which)
{
    int k = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    int scan = scans[k];
    float T[6];
    #pragma unroll
    for (int a = 0; a < 6; a++) T[a] = Ts[k * 6 + a];
    float4 *ptr;
    if (which == 0) {
        if (i >= p.counts[scan * 8 + 2]) return;
        ptr = (float4 *)p.lsharp_pts + (size_t)scan * p.cap_lsharp + i;
    } else {
        if (i >= p.counts[scan * 8 + 4]) return;
        ptr = (float4 *)p.lflat_pts + (size_t)scan * p.N + i;
    }
    float4 v = *ptr;
    float4 q = vlo_to_start(T, v, p.deskew, p.inv_period);
    float sx, cx, sy, cy, sz, cz;
    vlo_sincosf(T[0], sx, cx); vlo_sincosf(T[1], sy, cy); vlo_sincosf(T[2], sz, cz);
    float x = q.x, y = q.y, z = q.z;
    float x0 = x; x = cy * x0 + sy * z; z = cy * z - sy * x0;          // rotY(ry)
    float y0 = y; y = cx * y0 - sx * z; z = sx * y0 + cx * z;          // rotX(rx)
    x0 = x;       x = cz * x0 - sz * y; y = sz * x0 + cz * y;          // rotZ(rz)
    *ptr = make_float4(x + T[3], y + T[4], z + T[5], (float)(int)v.w);
}

static OdomParams make_params(vlo_handle *h)
{
    ScanBatchDev &sb = h->sb; const vlo_config &c = h->cfg;
    OdomParams p;
    p.sharp_pts = sb.sharp_pts; p.flat_pts = sb.flat_pts; p.cap_sharp = h->cap_sharp; p.cap_flat = h->cap_flat;
    p.lsharp_pts = sb.lsharp_pts; p.cap_lsharp = h->cap_lsharp; p.lflat_pts = sb.lflat_pts; p.N = c.max_points;
    p.lsharp_ring_start = sb.lsharp_ring_start;
    p.lflat_ring_start = sb.lflat_ring_start; p.counts = sb.counts; p.n_rings = c.n_rings;
    p.pair_last = h->pair_last; p.pair_cur = h->pair_cur; p.pair_T = h->pair_T; p.pair_state = h->pair_state; p.max_pairs = h->max_pairs;
    p.cidx = h->pair_cidx; p.sidx = h->pair_sidx; p.result = h->pair_result;
    p.gc = h->gs_corner; p.gsf = h->gs_surf; p.ss = h->segs;
    p.deskew = c.deskew; p.inv_period = 1.0f / c.scan_period; p.fwd_quirk = c.odom_forward_bound_quirk;
    p.max_iter = c.odom_max_iterations; p.degen_thr = c.odom_degen_eig; p.dT_abort = c.odom_delta_t_abort;
    p.dR_abort = c.odom_delta_r_abort; p.rot_thr = c.dopt_rot_threshold; p.trans_thr = c.dopt_trans_threshold;
    return p;
}

// builds the search indices of the target clouds of resident scans [first, first + count)
int vlo_build_scan_grids(vlo_handle *h, int first, int count)
{
    if (count <= 0) return VLO_OK;
    ScanBatchDev &sb = h->sb; const vlo_config &c = h->cfg;
#if K3_ONLINE_GRID
    h->scan_index_grid = count <= K3_THREAD_PAIRS;       // what the association of these scans will search
    if (h->scan_index_grid) {
        GridSource sc = {};
        sc.pts = sb.lsharp_pts; sc.pts_stride = (size_t)h->cap_lsharp;
        sc.n_dense = sb.counts; sc.n_dense_stride = 8; sc.n_dense_field = 2; sc.n_rings = c.n_rings;
        int rc = vlo_grid_build(h, h->gs_corner, sc, first, count, h->cap_lsharp); if (rc) return rc;
        GridSource ss = {};
        ss.pts = sb.lflat_pts; ss.pts_stride = (size_t)c.max_points;
        ss.n_dense = sb.counts; ss.n_dense_stride = 8; ss.n_dense_field = 4; ss.n_rings = c.n_rings;
        rc = vlo_grid_build(h, h->gs_surf, ss, first, count, c.max_points); if (rc) return rc;
    }
#endif
    SegBuildParams q;
    q.pts[0] = sb.lsharp_pts; q.stride[0] = (size_t)h->cap_lsharp; q.ring_start[0] = sb.lsharp_ring_start;
    q.pts[1] = sb.lflat_pts; q.stride[1] = (size_t)c.max_points; q.ring_start[1] = sb.lflat_ring_start;
    q.n_rings = c.n_rings; q.scan_first = first; q.ss = h->segs;
#if K3_ONLINE_GRID
    if (h->scan_index_grid) {
        VLO_PROF(h, ST_GRID_BUILD, (k3_partner_ranges<<<dim3(count, 2), SEGB_THREADS, 0, h->stream>>>(q)));
        h->launches += 1;
        VLO_CUDA(cudaGetLastError());
        return VLO_OK;
    }
#endif
    VLO_PROF(h, ST_GRID_BUILD, (k3_seg_build<<<dim3(count, 2), SEGB_THREADS, 0, h->stream>>>(q)));
    h->launches += 1;
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}

// transformToEnd of the target clouds (less sharp, less flat) of scans d_scans[0..n) with transforms d_T[n][6]
int vlo_launch_to_end(vlo_handle *h, const int *d_scans, const float *d_T, int n)
{
    OdomParams p = make_params(h);
    dim3 g0((h->cap_lsharp + 255) / 256, n), g1((h->cfg.max_points + 255) / 256, n);
    vlo_prof_begin(h, ST_TO_END);
    k3_to_end<<<g0, 256, 0, h->stream>>>(p, d_scans, d_T, n, 0);
    k3_to_end<<<g1, 256, 0, h->stream>>>(p, d_scans, d_T, n, 1);
    vlo_prof_end(h, ST_TO_END);
    h->launches += 2;
    h->grids_valid = 0;
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}

int vlo_launch_register_pairs(vlo_handle *h, int n_pairs, const float *d_seeds, const float *d_last_T, int only_grid_scan)
{
    OdomParams p = make_params(h);
    const vlo_config &c = h->cfg;
    k3_init_pairs<<<(n_pairs + 127) / 128, 128, 0, h->stream>>>(p, d_seeds, n_pairs);
    h->launches += 1;
    const size_t gn_smem = sizeof(float) * GN_THREADS * TSTRIDE;
    if (!h->k3_gn_configured) {                    // per handle = per device
        VLO_CUDA(cudaFuncSetAttribute(k3_gn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gn_smem));
        h->k3_gn_configured = 1;
    }
    if (d_last_T) {
        dim3 g0((h->cap_lsharp + 255) / 256, n_pairs), g1((c.max_points + 255) / 256, n_pairs);
        vlo_prof_begin(h, ST_TO_END);
        k3_to_end<<<g0, 256, 0, h->stream>>>(p, h->pair_last, d_last_T, n_pairs, 0);
        k3_to_end<<<g1, 256, 0, h->stream>>>(p, h->pair_last, d_last_T, n_pairs, 1);
        vlo_prof_end(h, ST_TO_END);
        h->launches += 2;
        h->grids_valid = 0;
    }
    if (only_grid_scan >= 0) { int rc = vlo_build_scan_grids(h, only_grid_scan, 1); if (rc) return rc; h->grids_valid = 1; }
    if (!h->grids_valid) { int rc = vlo_build_scan_grids(h, 0, h->sb.n_scans); if (rc) return rc; h->grids_valid = 1; }
    // association grid: every CTA owns a contiguous share of its pair's queries, one warp per query; a single pair (the
    // online tick) is spread over the whole machine, a batch gets as many CTAs per pair as keep every SM busy
    const int shares = (h->cap_sharp + h->cap_flat + K3A_WQ - 1) / K3A_WQ;
    if (c.odom_max_iterations > 5 * 64) { h->err = "odomMaxIterations: at most 320"; return VLO_ERR_INVALID_ARG; }
    if (!h->k3a_ctas) {                                  // persistent grid: every CTA slot of the device (per handle = per device)
        int occ = 0, sms = 0;
        VLO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_assoc, K3A_THREADS, 0));
        VLO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device));
        h->k3a_ctas = std::max(1, occ * sms);
    }
    const int ga = std::max(1, std::min(h->k3a_ctas, (n_pairs * shares + K3A_THREADS / 32 - 1) / (K3A_THREADS / 32)));
#if K3_ONLINE_GRID
    dim3 gw(std::max(1, std::min(((h->cap_sharp + h->cap_flat) * 32 + 255) / 256, (148 * 8 + n_pairs - 1) / n_pairs)), n_pairs);
#endif
    size_t trace_stride = (size_t)n_pairs * (h->cap_sharp * 2 + h->cap_flat * 3);
    for (int base = 0, round = 0; base < c.odom_max_iterations; base += 5, round++) {
#if K3_ONLINE_GRID
        if (h->scan_index_grid) VLO_PROF(h, ST_ASSOC, (k3_assoc_warp<<<gw, 256, 0, h->stream>>>(p, n_pairs)));
        else
#endif
        VLO_PROF(h, ST_ASSOC, (k3_assoc<<<ga, K3A_THREADS, 0, h->stream>>>(p, n_pairs, round, shares)));
        if (h->trace && round < 5) {
            int *dst = h->pair_trace + (size_t)round * trace_stride;
            VLO_CUDA(cudaMemcpyAsync(dst, h->pair_cidx, sizeof(int) * (size_t)n_pairs * h->cap_sharp * 2, cudaMemcpyDeviceToDevice, h->stream));
            VLO_CUDA(cudaMemcpyAsync(dst + (size_t)n_pairs * h->cap_sharp * 2, h->pair_sidx, sizeof(int) * (size_t)n_pairs * h->cap_flat * 3, cudaMemcpyDeviceToDevice, h->stream));
        }
        VLO_PROF(h, ST_GN, (k3_gn<<<n_pairs, GN_THREADS, gn_smem, h->stream>>>(p, base, 5)));
        h->launches += 2;
    }
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}
