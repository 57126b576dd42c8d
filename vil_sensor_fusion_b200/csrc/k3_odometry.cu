// K3/K5  scan-to-scan registration (LaserOdometry): association on the voxel-hash grids of the
// previous sweep + fused linearisation / deterministic reduction / on-device solve.
// Replaces BasicLaserOdometry::process of the `loam` nodelet (gtsam_fusion/launch/loam.launch:40-45;
// knobs loam_params.yaml:36-39); SURVEY.md Appendix A.4-A.7 is the algorithm, oracle/laser_odometry.c
// the frozen operation order (R1 three-level blocked summation, R2..R5).
//   k3_assoc : one warp per feature point: transformToStart, exact 1-NN, ring-constrained partner
//              search with upstream's forward/backward tie order, all on the grids (grid.cuh)
//   k3_gn    : one CTA per scan pair runs up to 5 Gauss-Newton iterations between associations:
//              thread per correspondence -> 28 products -> shared-memory transposed R1 reduction ->
//              warp 0 solves (QR), iteration 0 also runs the single-warp Jacobi degeneracy test.
// The host enqueues [k3_assoc, k3_gn] x ceil(maxIter/5) back to back; convergence is a device flag,
// so control never returns to the host inside a registration.
#include "grid.cuh"
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
    const int *pair_last, *pair_cur; float *pair_T; int *pair_state; int *cidx, *sidx; vlo_result *result;
    GridSet gc, gsf;                                     // corner / surf grids, grid index = scan index
    int deskew; float inv_period; int fwd_quirk;
    int max_iter; float degen_thr, dT_abort, dR_abort, rot_thr, trans_thr;
};

__device__ __forceinline__ float4 lflat_point(const OdomParams &p, int scan, int dense)
{
    return p.lflat_pts[(size_t)scan * p.N + dense];
}

__global__ void __launch_bounds__(256) k3_assoc(OdomParams p, int n_pairs)
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
                int ring = (int)(nn.tag[0] >> 24);
                FilterPartner f; f.ind = i1; f.ring_lo = ring - 2; f.ring_hi = ring + 2; f.skip_ring = ring;
                f.fwd_bound = p.fwd_quirk ? min(n_sharp, n_lc) : n_lc;
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
                int ring = (int)(nn.tag[0] >> 24);
                int fb = p.fwd_quirk ? min(n_flat, n_ls) : n_ls;
                FilterPartner f2; f2.ind = i1; f2.ring_lo = ring; f2.ring_hi = ring; f2.skip_ring = -1; f2.fwd_bound = fb;
                FilterPartner f3; f3.ind = i1; f3.ring_lo = ring - 2; f3.ring_hi = ring + 2; f3.skip_ring = ring; f3.fwd_bound = fb;
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

// Batch variant of k3_assoc (whole-bag mode): ONE THREAD per feature point.  The warp-per-query search above spends
// ~540 warp-instructions per query on its cooperative machinery (27 probes, prefix scan, merge rounds) -- fine for
// the latency of a single pair, wasteful when thousands of queries are waiting.  Here a thread runs the per-lane
// 27-cell search of grid.cuh for its own point, once per stage (nearest neighbour, partner(s)) through ONE inlined
// copy of the search.  Exactness: the 27-cell block around the query's cell contains every point closer than one cell
// edge, so a hit with d2 < (cell - slack)^2 is the global (d2, tie) minimum; a stage without such a hit (sparse
// regions, far partners) is redone with the exact warp-cooperative search, one such lane at a time.
struct FilterOdom {          // mode 0: plain nearest neighbour (tie = dense index); mode 1: FilterPartner's rules
    int mode, ind, ring_lo, ring_hi, skip_ring, fwd_bound;
    __device__ __forceinline__ bool operator()(unsigned tag, unsigned &tie) const
    {
        const int ring = (int)(tag >> 24), idx = (int)(tag & 0xFFFFFFu);
        if (mode == 0) { tie = (unsigned)idx; return true; }
        if (ring < ring_lo || ring > ring_hi || ring == skip_ring || idx == ind) return false;
        if (idx > ind) { if (idx >= fwd_bound) return false; tie = (unsigned)(idx - ind); }
        else tie = 0x40000000u + (unsigned)(ind - idx);
        return true;
    }
};

#define K3T_THREADS 128
#define K3_THREAD_PAIRS 4        // up to this many pairs per call keep the warp-per-query kernel (latency of the online tick)
__global__ void __launch_bounds__(K3T_THREADS) k3_assoc_thread(OdomParams p, int n_pairs)
{
    __shared__ int scratch[K3T_THREADS / 32][GRID_SCRATCH_INTS];
    const int pair = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (p.pair_state[pair * 4 + 0]) return;                          // converged
    const int last = p.pair_last[pair], cur = p.pair_cur[pair];
    const int n_sharp = p.counts[cur * 8 + 1], n_flat = p.counts[cur * 8 + 3];
    const int n_lc = p.counts[last * 8 + 2], n_ls = p.counts[last * 8 + 4];
    if (!(n_lc > 10 && n_ls > 100)) return;
    float T[6];
    #pragma unroll
    for (int a = 0; a < 6; a++) T[a] = p.pair_T[pair * 6 + a];
    const int total = n_sharp + n_flat, n_warps = gridDim.x * (K3T_THREADS / 32);
    for (int base = (blockIdx.x * (K3T_THREADS / 32) + warp) * 32; base < total; base += n_warps * 32) {   // warp-uniform
        const int w = base + lane;
        const bool valid = w < total, sharp = w < n_sharp;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) q = vlo_to_start(T, sharp ? p.sharp_pts[(size_t)cur * p.cap_sharp + w] : p.flat_pts[(size_t)cur * p.cap_flat + (w - n_sharp)],
                                    p.deskew, p.inv_period);
        const GridSet gs = sharp ? p.gc : p.gsf;
        const float edge = gs.cell - 2e-3f * gs.cell;
        const float dfast = fminf(25.0f, edge * edge);
        const int fb_bound = sharp ? (p.fwd_quirk ? min(n_sharp, n_lc) : n_lc) : (p.fwd_quirk ? min(n_flat, n_ls) : n_ls);
        int i1 = -1, i2 = -1, i3 = -1, ring = 0;
        #pragma unroll 1
        for (int stage = 0; stage < 3; stage++) {
            const bool act = valid && (stage == 0 || i1 >= 0) && (stage < 2 || !sharp);
            FilterOdom f;
            f.mode = stage == 0 ? 0 : 1; f.ind = i1; f.fwd_bound = fb_bound;
            if (stage == 1 && !sharp) { f.ring_lo = ring; f.ring_hi = ring; f.skip_ring = -1; }      // same-ring partner of a flat point
            else { f.ring_lo = ring - 2; f.ring_hi = ring + 2; f.skip_ring = ring; }
            unsigned tag = GRID_NOTAG;
            if (act) {
                TopKT<1> best;
                grid_search_thread27(gs, last, q.x, q.y, q.z, dfast, f, best);
                tag = best.tag[0];
            }
            // stages the 27-cell block could not settle: exact warp-cooperative search, one lane's query at a time
            unsigned redo = __ballot_sync(0xffffffffu, act && tag == GRID_NOTAG && dfast < 25.0f);
            while (redo) {
                const int src = __ffs(redo) - 1;
                redo &= redo - 1u;
                FilterOdom g;
                g.mode = __shfl_sync(0xffffffffu, f.mode, src); g.ind = __shfl_sync(0xffffffffu, f.ind, src);
                g.ring_lo = __shfl_sync(0xffffffffu, f.ring_lo, src); g.ring_hi = __shfl_sync(0xffffffffu, f.ring_hi, src);
                g.skip_ring = __shfl_sync(0xffffffffu, f.skip_ring, src); g.fwd_bound = __shfl_sync(0xffffffffu, f.fwd_bound, src);
                const float bx = __shfl_sync(0xffffffffu, q.x, src), by = __shfl_sync(0xffffffffu, q.y, src), bz = __shfl_sync(0xffffffffu, q.z, src);
                const bool bsharp = __shfl_sync(0xffffffffu, (int)sharp, src) != 0;
                TopK<1> r;
                grid_search<1>(bsharp ? p.gc : p.gsf, last, bx, by, bz, 25.0f, g, r, lane, scratch[warp]);
                if (lane == src) tag = r.tag[0];
            }
            const int idx = (act && tag != GRID_NOTAG) ? (int)(tag & 0xFFFFFFu) : -1;
            if (stage == 0) { i1 = idx; ring = (int)(tag >> 24); }
            else if (stage == 1) i2 = idx;
            else i3 = idx;
        }
        if (valid) {
            if (sharp) { int *o = p.cidx + ((size_t)pair * p.cap_sharp + w) * 2; o[0] = i1; o[1] = i2; }
            else { int *o = p.sidx + ((size_t)pair * p.cap_flat + (w - n_sharp)) * 3; o[0] = i1; o[1] = i2; o[2] = i3; }
        }
    }
}

#define GN_THREADS 256
#define TSTRIDE 29     // padded row of the transposed term buffer

// R1 block reduction of one chunk of 256 queries held in smem terms[256][TSTRIDE]:
// 224 threads sum 32 consecutive queries each (level 1), then 28 threads fold the 8 level-1 sums into
// the running level-2 accumulator; level 2 closes into level 3 every 1024 queries.
__device__ __forceinline__ void r1_chunk(float *terms, float *l1buf, float *l2acc, float *l3acc, int chunk, int n_chunks, int q_total, int tid)
{
    __syncthreads();
    if (tid < 8 * VLO_NTERM) {
        int g = tid / VLO_NTERM, e = tid % VLO_NTERM;
        float l1 = 0.0f;
        const float *src = terms + (size_t)(g * 32) * TSTRIDE + e;
        #pragma unroll 8
        for (int k = 0; k < 32; k++) l1 = l1 + src[(size_t)k * TSTRIDE];
        l1buf[g * VLO_NTERM + e] = l1;
    }
    __syncthreads();
    if (tid < VLO_NTERM) {
        float l2 = l2acc[tid];
        for (int g = 0; g < 8; g++) if (chunk * 256 + g * 32 < q_total) l2 = l2 + l1buf[g * VLO_NTERM + tid];
        if ((chunk & 3) == 3 || chunk == n_chunks - 1) { l3acc[tid] = l3acc[tid] + l2; l2 = 0.0f; }
        l2acc[tid] = l2;
    }
}

__global__ void __launch_bounds__(GN_THREADS) k3_gn(OdomParams p, int iter_base, int n_iters)
{
    __shared__ float terms[GN_THREADS * TSTRIDE];
    __shared__ float l1buf[8 * VLO_NTERM];
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
    const int n_chunks = (q_total + 255) / 256;
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
            int i = chunk * 256 + tid;
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
    p.pair_last = h->pair_last; p.pair_cur = h->pair_cur; p.pair_T = h->pair_T; p.pair_state = h->pair_state;
    p.cidx = h->pair_cidx; p.sidx = h->pair_sidx; p.result = h->pair_result;
    p.gc = h->gs_corner; p.gsf = h->gs_surf;
    p.deskew = c.deskew; p.inv_period = 1.0f / c.scan_period; p.fwd_quirk = c.odom_forward_bound_quirk;
    p.max_iter = c.odom_max_iterations; p.degen_thr = c.odom_degen_eig; p.dT_abort = c.odom_delta_t_abort;
    p.dR_abort = c.odom_delta_r_abort; p.rot_thr = c.dopt_rot_threshold; p.trans_thr = c.dopt_trans_threshold;
    return p;
}

// builds the corner / surf grids of resident scans [first, first + count) (grid index = scan index)
int vlo_build_scan_grids(vlo_handle *h, int first, int count)
{
    ScanBatchDev &sb = h->sb; const vlo_config &c = h->cfg;
    GridSource sc = {};
    sc.pts = sb.lsharp_pts; sc.pts_stride = (size_t)h->cap_lsharp; sc.ring_off = nullptr; sc.ring_off_stride = 0;
    sc.ring_cnt = nullptr; sc.ring_cnt_stride = 0; sc.dense_start = nullptr; sc.dense_start_stride = 0;
    sc.n_dense = sb.counts; sc.n_dense_stride = 8; sc.n_dense_field = 2; sc.n_rings = c.n_rings; sc.grid_scan = nullptr;
    int rc = vlo_grid_build(h, h->gs_corner, sc, first, count, h->cap_lsharp); if (rc) return rc;
    GridSource ss = {};
    ss.pts = sb.lflat_pts; ss.pts_stride = (size_t)c.max_points; ss.ring_off = nullptr; ss.ring_off_stride = 0;
    ss.ring_cnt = nullptr; ss.ring_cnt_stride = 0; ss.dense_start = nullptr; ss.dense_start_stride = 0;
    ss.n_dense = sb.counts; ss.n_dense_stride = 8; ss.n_dense_field = 4; ss.n_rings = c.n_rings; ss.grid_scan = nullptr;
    rc = vlo_grid_build(h, h->gs_surf, ss, first, count, c.max_points); if (rc) return rc;
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
    // persistent association grid: enough CTAs to fill the machine, warps stride over the queries
    int n_warps = h->cap_sharp + h->cap_flat;
    int ctas = (n_warps * 32 + 255) / 256;
    int fill = (148 * 8 + n_pairs - 1) / n_pairs;
    dim3 ga(std::max(1, std::min(ctas, fill)), n_pairs);
    // batches: one thread per query (k3_assoc_thread), never more threads than queries
    dim3 gt(std::max(1, (h->cap_sharp + h->cap_flat + K3T_THREADS - 1) / K3T_THREADS), n_pairs);
    size_t trace_stride = (size_t)n_pairs * (h->cap_sharp * 2 + h->cap_flat * 3);
    for (int base = 0, round = 0; base < c.odom_max_iterations; base += 5, round++) {
        if (n_pairs > K3_THREAD_PAIRS) VLO_PROF(h, ST_ASSOC, (k3_assoc_thread<<<gt, K3T_THREADS, 0, h->stream>>>(p, n_pairs)));
        else VLO_PROF(h, ST_ASSOC, (k3_assoc<<<ga, 256, 0, h->stream>>>(p, n_pairs)));
        if (h->trace && round < 5) {
            int *dst = h->pair_trace + (size_t)round * trace_stride;
            VLO_CUDA(cudaMemcpyAsync(dst, h->pair_cidx, sizeof(int) * (size_t)n_pairs * h->cap_sharp * 2, cudaMemcpyDeviceToDevice, h->stream));
            VLO_CUDA(cudaMemcpyAsync(dst + (size_t)n_pairs * h->cap_sharp * 2, h->pair_sidx, sizeof(int) * (size_t)n_pairs * h->cap_flat * 3, cudaMemcpyDeviceToDevice, h->stream));
        }
        VLO_PROF(h, ST_GN, (k3_gn<<<n_pairs, GN_THREADS, 0, h->stream>>>(p, base, 5)));
        h->launches += 2;
    }
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}
