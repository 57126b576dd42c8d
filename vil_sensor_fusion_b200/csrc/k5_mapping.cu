// K5  scan-to-map registration (LaserMapping) against a device-resident voxel-hash map.
// Replaces BasicLaserMapping::optimizeTransformTobeMapped of the `loam` nodelet
// (gtsam_fusion/launch/loam.launch:47-52; knobs loam_params.yaml:44-46,53); SURVEY.md Appendix A.8 is
// the algorithm, oracle/laser_mapping.c the frozen operation order.  Per Gauss-Newton iteration:
//   k5_tile<MODE>  one thread per feature point of a 32-point warp tile: pointAssociateToMap + exact 5-NN (d2 < 1) on
//            the map grid (grid.cuh grid_search_thread27: per-lane 27-cell work mask, nearest-first, box pruning), then
//            3x3 covariance eigen (corner) / 5x3 least-squares plane (surface), residual, Jacobian row, 28 products;
//            level-1 sums of the R1 reduction per tile
//   solve    levels 2/3 of R1 in fixed order, QR solve, (iteration 0) Jacobi degeneracy test + remapping, pose update,
//            convergence, result record -- by a CTA (k5_solve_slot) or by a single warp (k5_solve_slot_warp)
// Two launch strategies over the same device functions:
//   batches      k5_assoc (association) + k5_lin (linearisation; the warp completing a slot's last tile solves it) per
//                iteration, persistent grids over one flat ticketed tile list of the unconverged slots
//   online tick  k5_register_coop: everything in ONE cooperative launch with grid syncs between the phases -- the
//                Gauss-Newton loop lives on the device and ends there
#include "grid.cuh"
#include "dense6.cuh"
#include "map_lin.cuh"
#include <algorithm>
#ifndef VLO_HOST_EMULATION              // tests/host/cuda_emul.h supplies cooperative_groups::this_grid() for the CPU emulation
#include <cooperative_groups.h>
#endif

struct MapParams {
    const float4 *lsharp_pts; int cap_lsharp; const float4 *lflat_pts; int N;
    const int *counts; int n_rings;
    const int *scans;            // [n] resident scan index per slot
    float *T; int *state; vlo_result *result;     // per slot
    int *idx5; int qcap;         // [n][qcap][5]
    float *partials; int pcap;   // [n][pcap][28] level-1 sums
    int *ncorr;                  // [n][2]
    int *done;                   // completion counters, convergence stamps, converged count, tile tickets (layout: see k5_build_list)
    GridSet gm0, gm1; int rho0, rho1; const float4 *map0, *map1; const int *map_n;
    int max_iter; float degen_thr, dT_abort, dR_abort, rot_thr, trans_thr;
};

__device__ __forceinline__ float4 map_query_point(const MapParams &p, int scan, int i, int n_ls, bool &corner)
{
    corner = i < n_ls;
    if (corner) return p.lsharp_pts[(size_t)scan * p.cap_lsharp + i];
    return p.lflat_pts[(size_t)scan * p.N + (i - n_ls)];
}

#define KNN_THREADS 128
#define AL_WARPS (KNN_THREADS / 32)
#define VLO_COOP_MAX_SLOTS 4      // registrations of up to this many scans run as one cooperative launch

#define LSTRIDE 29

// ---- association + linearisation of one warp tile ------------------------------------------------
// Warp tile t of slot k owns feature points [32 t, 32 t + 32) of the scan (corners first, then surface
// points); a thread does association AND linearisation for its point -- exact 5-NN on the map grid,
// line / plane fit, residual, Jacobian row, 28 products -- so the neighbour coordinates never leave the
// SM between the two; the warp then closes level 1 of the R1 sum (its 32 consecutive points, sequential
// order) through a per-warp shared-memory transpose.  From the second iteration on, the previous
// iteration's neighbours (still in idx5) seen from the new pose bound the 5th-neighbour distance, so the
// cell walk starts with a tight pruning radius.
// MODE 0: association + linearisation fused (the cooperative online-tick kernel); MODE 1: association only (writes idx5);
// MODE 2: linearisation only (reads idx5).  The batch path runs 1 and 2 as separate kernels: each then has a small
// instruction footprint and register budget (the fused kernel was 150 KB of SASS at 20 warps per SM and stalled on
// instruction fetch as soon as its warps ran out of step); the arithmetic is the same, statement for statement.
template <int MODE>
__device__ __forceinline__ void k5_tile(const MapParams &p, int k, int tile, int it, float *wterms, int lane)
{
    const int scan = p.scans[k];
    const int n_ls = p.counts[scan * 8 + 2], n_lf = p.counts[scan * 8 + 4];
    const int q_total = n_ls + n_lf;
    float T[6], trig[6];
    {
        float sv = 0.0f, cv = 0.0f;
        if (lane < 3) vlo_sincosf(__ldcg(p.T + k * 6 + lane), sv, cv);
        #pragma unroll
        for (int a = 0; a < 3; a++) { trig[2 * a] = __shfl_sync(0xffffffffu, sv, a); trig[2 * a + 1] = __shfl_sync(0xffffffffu, cv, a); }
        #pragma unroll
        for (int a = 0; a < 6; a++) T[a] = __ldcg(p.T + k * 6 + a);
    }
    const int i = tile * 32 + lane;
    float t[VLO_NTERM];
    #pragma unroll
    for (int e = 0; e < VLO_NTERM; e++) t[e] = 0.0f;
    int my_e = 0, my_p = 0;
    if (i < q_total) {
        bool corner;
        const float4 ori = map_query_point(p, scan, i, n_ls, corner);
        const float4 sel = to_map(T, trig, ori);
        const float4 *map = corner ? p.map0 : p.map1;
        int *o = p.idx5 + ((size_t)k * p.qcap + i) * 5;
        bool ok;
        int nbi[5];
        if constexpr (MODE != 2) {
            float bound = -1.0f;
            if (it > 0 && o[4] >= 0) {
                bound = 0.0f;
                #pragma unroll
                for (int j = 0; j < 5; j++) {
                    const float4 m = map[o[j]];
                    const float ddx = m.x - sel.x, ddy = m.y - sel.y, ddz = m.z - sel.z;
                    bound = fmaxf(bound, (ddx * ddx + ddy * ddy) + ddz * ddz);
                }
            }
            TopKI<5> best;
            // one search call for both clouds (per-lane grid descriptor): a second inlined copy of the search doubled
            // the hot loop's instruction footprint
            GridSet gm = corner ? p.gm0 : p.gm1;
            const int rho = corner ? p.rho0 : p.rho1;
            if (rho == 1) grid_search_thread27(gm, 0, sel.x, sel.y, sel.z, 1.0f, FilterAll(), best, bound);
            else grid_search_thread(gm, 0, sel.x, sel.y, sel.z, 1.0f, rho, FilterAll(), best, bound);
            ok = best.valid(4);
            #pragma unroll
            for (int j = 0; j < 5; j++) { nbi[j] = ok ? best.index(j) : -1; o[j] = nbi[j]; }
        } else {
            #pragma unroll
            for (int j = 0; j < 5; j++) nbi[j] = o[j];
            ok = nbi[4] >= 0;
        }
        if constexpr (MODE == 1) return;
        if (ok) {
            float4 nb[5];
            #pragma unroll
            for (int j = 0; j < 5; j++) nb[j] = map[nbi[j]];
            float coeff[4];
            bool keep = corner ? map_edge_coeff(sel, nb, coeff) : map_plane_coeff(sel, nb, coeff);
            if (keep) {
                if (corner) my_e = 1; else my_p = 1;
                float srx = trig[0], crx = trig[1], sry = trig[2], cry = trig[3], srz = trig[4], crz = trig[5];
                float x = ori.x, y = ori.y, z = ori.z, cx_ = coeff[0], cy_ = coeff[1], cz_ = coeff[2];
                float row[6];
                row[0] = (crx * sry * srz * x + crx * crz * sry * y - srx * sry * z) * cx_
                       + (-srx * srz * x - crz * srx * y - crx * z) * cy_
                       + (crx * cry * srz * x + crx * cry * crz * y - cry * srx * z) * cz_;
                row[1] = ((cry * srx * srz - crz * sry) * x + (sry * srz + cry * crz * srx) * y + crx * cry * z) * cx_
                       + ((-cry * crz - srx * sry * srz) * x + (cry * srz - crz * srx * sry) * y - crx * sry * z) * cz_;
                row[2] = ((crz * srx * sry - cry * srz) * x + (-cry * crz - srx * sry * srz) * y) * cx_
                       + (crx * crz * x - crx * srz * y) * cy_
                       + ((sry * srz + cry * crz * srx) * x + (crz * sry - cry * srx * srz) * y) * cz_;
                row[3] = cx_; row[4] = cy_; row[5] = cz_;
                float bval = -coeff[3];
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
    }
    if constexpr (MODE == 1) return;
    __syncwarp();                       // the previous tile's column sums are done with wterms
    #pragma unroll
    for (int e = 0; e < VLO_NTERM; e++) wterms[lane * LSTRIDE + e] = t[e];
    const int ne = (int)__reduce_add_sync(0xffffffffu, my_e), np = (int)__reduce_add_sync(0xffffffffu, my_p);
    __syncwarp();
    if (lane < VLO_NTERM) {
        float l1 = 0.0f;
        #pragma unroll 8
        for (int q = 0; q < 32; q++) l1 = l1 + wterms[q * LSTRIDE + lane];
        p.partials[((size_t)k * p.pcap + tile) * VLO_NTERM + lane] = l1;
    }
    if (lane == 0 && (ne | np)) { atomicAdd(&p.ncorr[k * 2], ne); atomicAdd(&p.ncorr[k * 2 + 1], np); }
}

// ---- levels 2/3 of R1, solve, degeneracy, pose update of one slot (whole CTA) ----------------------
struct SolveSmem { GnScratch S; float l2[64 * VLO_NTERM]; };

__device__ __forceinline__ void k5_solve_slot(const MapParams &p, int k, int it, SolveSmem &M)
{
    GnScratch &S = M.S; float *l2 = M.l2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    int *state = p.state + k * 4;
    const int scan = p.scans[k];
    vlo_result *res = p.result + k;
    const int q_total = p.counts[scan * 8 + 2] + p.counts[scan * 8 + 4];
    if (!(p.map_n[2] > 10 && p.map_n[4] > 100)) {
        if (tid == 0) { state[0] = 1; state[1] = 0; res->iterations = 0; res->status = VLO_SOFT_TOO_FEW_CORR; }
        return;
    }
    const int n_edge = __ldcg(p.ncorr + k * 2), n_plane = __ldcg(p.ncorr + k * 2 + 1);
    __syncthreads();
    if (tid == 0) { p.ncorr[k * 2] = 0; p.ncorr[k * 2 + 1] = 0; state[1] = it + 1; res->iterations = it + 1; }
    if (n_edge + n_plane < 50) return;            // upstream `continue`
    if (tid < 6) S.T[tid] = __ldcg(p.T + k * 6 + tid);
    if (tid < 36 && it > 0) S.P[tid] = __ldcg(&res->P[tid]);
    if (tid == 0) { S.is_degenerate = __ldcg(state + 2); S.converged = 0; S.n_edge = n_edge; S.n_plane = n_plane; }
    // R1 levels 2 and 3 over the level-1 sums
    const int n_l1 = (q_total + 31) / 32, n_l2 = (n_l1 + 31) / 32;
    float l3 = 0.0f;
    for (int base = 0; base < n_l2; base += 64) {
        __syncthreads();
        for (int task = tid; task < 64 * VLO_NTERM; task += nthr) {
            int b2 = base + task / VLO_NTERM, e = task % VLO_NTERM;
            if (b2 < n_l2) {
                float acc = 0.0f;
                int lo = b2 * 32, hi = min(n_l1, lo + 32);
                const float *src = p.partials + ((size_t)k * p.pcap + lo) * VLO_NTERM + e;
                for (int q = lo; q < hi; q++, src += VLO_NTERM) acc = acc + __ldcg(src);
                l2[(task / VLO_NTERM) * VLO_NTERM + e] = acc;
            }
        }
        __syncthreads();
        if (tid < VLO_NTERM) {
            int cnt = min(64, n_l2 - base);
            for (int b = 0; b < cnt; b++) l3 = l3 + l2[b * VLO_NTERM + tid];
        }
    }
    if (tid < VLO_NTERM) S.total[tid] = l3;
    __syncthreads();
    if (warp == 0) vlo_gn_update_warp(S, it, p.degen_thr, p.dT_abort, p.dR_abort, lane);
    __syncthreads();
    if (tid < 6) { p.T[k * 6 + tid] = S.T[tid]; res->transform[tid] = S.T[tid]; }
    if (it == 0) {
        if (tid < 36) res->P[tid] = S.P[tid];
        if (tid < 6) res->eig[tid] = S.eval[tid];
    }
    if (tid == 0) {
        state[0] = S.converged; state[2] = S.is_degenerate; state[3] = VLO_OK;
        res->is_degenerate = S.is_degenerate; res->status = VLO_OK;
        res->n_corr_edge = n_edge; res->n_corr_plane = n_plane;
        vlo_finish_result(S, p.rot_thr, p.trans_thr, res);
    }
    __syncthreads();
}

// ---- the whole registration of a batch in ONE cooperative launch ------------------------------------
// Every Gauss-Newton iteration is [association + linearisation of all unconverged slots] -> grid sync ->
// [solve per slot] -> grid sync; the loop ends on the device when every slot has converged (or at
// mapMaxIterations), so neither the host nor empty launches sit between iterations (north star (4)).
// Work distribution: the warp tiles of all unconverged slots form one flat list (prefix sums rebuilt by
// every CTA per iteration) that the persistent warps stride over, so late iterations with few active
// slots still use the whole machine.
namespace cg = cooperative_groups;

// Throughput path (large batches): one launch per Gauss-Newton iteration over ONE flat list of the warp tiles of
// all unconverged slots.  ncu r01c of the previous (slot, CTA-strided) launch shape: every warp owned a single tile,
// per-tile cost varies 3x, and warps sat at the CTA's closing barrier for 11 % of the stall samples with 30 % of the
// warp slots occupied.  Here
//   * every CTA rebuilds the same compacted list (slot, tile prefix) from the convergence stamps,
//   * a warp takes its first tile by position and every further one by global ticket (drawn before the current tile
//     is worked on), so no warp idles while tiles are left anywhere in the batch,
//   * the warp that completes a slot's last tile (per-slot completion counter) closes R1 levels 2/3, solves, tests
//     degeneracy and updates the pose of that slot on its own -- a warp-level solve, no CTA barrier anywhere.
// Convergence is a stamp (iteration at which the slot converged) rather than a flag: a slot solved early in launch
// `it` must stay in the list of CTAs that start later in the same launch, or their ticket numbering would differ.
struct SolveWarpSmem { GnScratch S; float l2[8 * VLO_NTERM]; };
#define K5_NOT_CONVERGED 0x7fffffff

__device__ __noinline__ void k5_solve_slot_warp(const MapParams &p, int k, int it, SolveWarpSmem &M, int lane, int n)
{
    GnScratch &S = M.S;
    int *state = p.state + k * 4;
    const int scan = p.scans[k];
    vlo_result *res = p.result + k;
    const int q_total = p.counts[scan * 8 + 2] + p.counts[scan * 8 + 4];
    if (!(p.map_n[2] > 10 && p.map_n[4] > 100)) {
        if (lane == 0) { state[0] = 1; state[1] = 0; res->iterations = 0; res->status = VLO_SOFT_TOO_FEW_CORR; p.done[2 * k + 1] = it; atomicAdd(&p.done[2 * n], 1); }
        return;
    }
    const int n_edge = __ldcg(p.ncorr + k * 2), n_plane = __ldcg(p.ncorr + k * 2 + 1);
    __syncwarp();
    if (lane == 0) { p.ncorr[k * 2] = 0; p.ncorr[k * 2 + 1] = 0; state[1] = it + 1; res->iterations = it + 1; }
    if (n_edge + n_plane < 50) return;            // upstream `continue`
    if (lane < 6) S.T[lane] = __ldcg(p.T + k * 6 + lane);
    if (it > 0) for (int e = lane; e < 36; e += 32) S.P[e] = __ldcg(&res->P[e]);
    if (lane == 0) { S.is_degenerate = __ldcg(state + 2); S.converged = 0; S.n_edge = n_edge; S.n_plane = n_plane; }
    // R1 levels 2 and 3 over the level-1 sums, same order as k5_solve_slot
    const int n_l1 = (q_total + 31) / 32, n_l2 = (n_l1 + 31) / 32;
    float l3 = 0.0f;
    for (int base = 0; base < n_l2; base += 8) {
        __syncwarp();
        for (int task = lane; task < 8 * VLO_NTERM; task += 32) {
            const int b2 = base + task / VLO_NTERM, e = task % VLO_NTERM;
            if (b2 < n_l2) {
                float acc = 0.0f;
                const int lo = b2 * 32, hi = min(n_l1, lo + 32);
                const float *src = p.partials + ((size_t)k * p.pcap + lo) * VLO_NTERM + e;
                for (int q = lo; q < hi; q++, src += VLO_NTERM) acc = acc + __ldcg(src);
                M.l2[task] = acc;
            }
        }
        __syncwarp();
        if (lane < VLO_NTERM) {
            const int cnt = min(8, n_l2 - base);
            for (int b = 0; b < cnt; b++) l3 = l3 + M.l2[b * VLO_NTERM + lane];
        }
    }
    if (lane < VLO_NTERM) S.total[lane] = l3;
    __syncwarp();
    vlo_gn_update_warp(S, it, p.degen_thr, p.dT_abort, p.dR_abort, lane);
    if (lane < 6) { p.T[k * 6 + lane] = S.T[lane]; res->transform[lane] = S.T[lane]; }
    if (it == 0) {
        for (int e = lane; e < 36; e += 32) res->P[e] = S.P[e];
        if (lane < 6) res->eig[lane] = S.eval[lane];
    }
    if (lane == 0) {
        state[0] = S.converged; state[2] = S.is_degenerate; state[3] = VLO_OK;
        if (S.converged) { p.done[2 * k + 1] = it; atomicAdd(&p.done[2 * n], 1); }
        res->is_degenerate = S.is_degenerate; res->status = VLO_OK;
        res->n_corr_edge = n_edge; res->n_corr_plane = n_plane;
        vlo_finish_result(S, p.rot_thr, p.trans_thr, res);
    }
    __syncwarp();
}

// p.done: [n][2] = {tiles of the slot completed in this launch, iteration at which the slot converged}, then
// [2 n] = number of slots that have converged (every later launch returns at once when it reaches n), then one ticket
// counter per launch at [2 n + 1 + 2 it + phase]
struct ListSmem { int warp_tot[AL_WARPS], warp_act[AL_WARPS], n_active; };

// flat work list: the unconverged slots (order-preserving) with the prefix sums of their tile counts; returns the
// number of active slots (the same in every CTA of the launch)
__device__ __forceinline__ int k5_build_list(const MapParams &p, int it, int n, int *s_slot, int *s_pref, ListSmem &L)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { L.n_active = 0; s_pref[0] = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += KNN_THREADS) {
        const int k = base + tid;
        int tiles = 0; bool act = false;
        if (k < n && __ldcg(p.done + 2 * k + 1) >= it) {
            const int scan = p.scans[k];
            tiles = max(1, (p.counts[scan * 8 + 2] + p.counts[scan * 8 + 4] + 31) >> 5);   // an empty slot still gets solved
            act = true;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, act);
        int inc = tiles;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
        if (lane == 31) L.warp_tot[warp] = inc;
        if (lane == 0) L.warp_act[warp] = __popc(bal);
        __syncthreads();
        int a0 = L.n_active, t0 = s_pref[a0];
        for (int w = 0; w < warp; w++) { a0 += L.warp_act[w]; t0 += L.warp_tot[w]; }
        if (act) {
            const int pos = a0 + __popc(bal & ((1u << lane) - 1u));
            s_slot[pos] = k;
            s_pref[pos + 1] = t0 + inc;
        }
        __syncthreads();
        if (tid == 0) { int a = 0; for (int w = 0; w < AL_WARPS; w++) a += L.warp_act[w]; L.n_active += a; }
        __syncthreads();
    }
    return L.n_active;
}

// association of every unconverged slot's feature points (exact 5-NN, indices to idx5)
#ifdef K5A_MINB
__global__ void __launch_bounds__(KNN_THREADS, K5A_MINB) k5_assoc(MapParams p, int it, int n)
#else
__global__ void __launch_bounds__(KNN_THREADS) k5_assoc(MapParams p, int it, int n)
#endif
{
    VLO_DYN_SMEM_INT(s_dyn);                       // [n] active slots, [n + 1] tile prefix
    __shared__ ListSmem L;
    if (__ldcg(p.done + 2 * n) >= n) return;       // everything converged in an earlier launch
    if (!(p.map_n[2] > 10 && p.map_n[4] > 100)) return;
    int *s_slot = s_dyn, *s_pref = s_dyn + n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_active = k5_build_list(p, it, n, s_slot, s_pref, L);
    if (n_active == 0) return;
    const int total = s_pref[n_active];
    const int n_static = gridDim.x * AL_WARPS;
    int g = blockIdx.x * AL_WARPS + warp;
    while (g < total) {
        int nxt = 0;
        if (lane == 0) nxt = n_static + atomicAdd(&p.done[2 * n + 1 + 2 * it], 1);
        int lo = 0, hi = n_active;                  // slot j with pref[j] <= g < pref[j + 1]
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pref[mid] <= g) lo = mid; else hi = mid; }
        k5_tile<1>(p, s_slot[lo], g - s_pref[lo], it, nullptr, lane);
        g = __shfl_sync(0xffffffffu, nxt, 0);
    }
}

// linearisation + level-1 sums of every unconverged slot; the warp that completes a slot's last tile solves the slot
// (8 CTAs per SM at 64 registers measured the same as 6 at 80: 0.1528 vs 0.1529 ms per launch -- not occupancy-bound)
__global__ void __launch_bounds__(KNN_THREADS, 6) k5_lin(MapParams p, int it, int n)
{
    VLO_DYN_SMEM_INT(s_dyn);
    __shared__ float terms[AL_WARPS][32 * LSTRIDE];
    __shared__ SolveWarpSmem M[AL_WARPS];
    __shared__ ListSmem L;
    if (__ldcg(p.done + 2 * n) >= n) return;
    int *s_slot = s_dyn, *s_pref = s_dyn + n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_active = k5_build_list(p, it, n, s_slot, s_pref, L);
    if (n_active == 0) return;
    const int total = s_pref[n_active];
    const bool map_ok = p.map_n[2] > 10 && p.map_n[4] > 100;
    const int n_static = gridDim.x * AL_WARPS;
    int g = blockIdx.x * AL_WARPS + warp;
    while (g < total) {
        int nxt = 0;
        if (lane == 0) nxt = n_static + atomicAdd(&p.done[2 * n + 2 + 2 * it], 1);
        int lo = 0, hi = n_active;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pref[mid] <= g) lo = mid; else hi = mid; }
        const int k = s_slot[lo], tile = g - s_pref[lo], n_tiles = s_pref[lo + 1] - s_pref[lo];
        if (map_ok) k5_tile<2>(p, k, tile, it, terms[warp], lane);
        __threadfence();                            // this tile's level-1 sums and counts before the completion count
        __syncwarp();
        int last = 0;
        if (lane == 0) last = atomicAdd(&p.done[2 * k], 1) == n_tiles - 1;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if (lane == 0) p.done[2 * k] = 0;
            __threadfence();
            k5_solve_slot_warp(p, k, it, M[warp], lane, n);
        }
        g = __shfl_sync(0xffffffffu, nxt, 0);
    }
}

// Latency path (online tick, a few slots): the whole registration in ONE cooperative launch.

__global__ void __launch_bounds__(KNN_THREADS, 5) k5_register_coop(MapParams p, int n)
{
    VLO_DYN_SMEM_INT(s_dyn);                       // [n] active slots, [n + 1] tile prefix
    __shared__ float terms[AL_WARPS][32 * LSTRIDE];
    __shared__ SolveSmem M;
    __shared__ int s_warp_tot[AL_WARPS], s_n_active;
    cg::grid_group grid = cg::this_grid();
    int *s_slot = s_dyn, *s_pref = s_dyn + n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int it = 0; it < p.max_iter; it++) {
        // flat work list: compact the unconverged slots (order-preserving) with their tile counts
        if (tid == 0) { s_n_active = 0; s_pref[0] = 0; }
        __syncthreads();
        for (int base = 0; base < n; base += KNN_THREADS) {
            const int k = base + tid;
            int tiles = 0; bool act = false;
            if (k < n && !__ldcg(p.state + k * 4)) {
                const int scan = p.scans[k];
                tiles = (p.counts[scan * 8 + 2] + p.counts[scan * 8 + 4] + 31) >> 5;
                act = true;
            }
            // block-wide exclusive scans of (act, tiles)
            const unsigned bal = __ballot_sync(0xffffffffu, act);
            int inc = tiles;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
            if (lane == 31) s_warp_tot[warp] = inc;
            __shared__ int s_warp_act[AL_WARPS];
            if (lane == 0) s_warp_act[warp] = __popc(bal);
            __syncthreads();
            int a0 = s_n_active, t0 = s_pref[a0];
            for (int w = 0; w < warp; w++) { a0 += s_warp_act[w]; t0 += s_warp_tot[w]; }
            if (act) {
                const int pos = a0 + __popc(bal & ((1u << lane) - 1u));
                s_slot[pos] = k;
                s_pref[pos + 1] = t0 + inc;
            }
            __syncthreads();
            if (tid == 0) { int a = 0; for (int w = 0; w < AL_WARPS; w++) a += s_warp_act[w]; s_n_active += a; }
            __syncthreads();
        }
        const int n_active = s_n_active;
        if (n_active == 0) break;                   // same decision in every CTA: state is grid-uniform here
        const int total = s_pref[n_active];
        for (int g = blockIdx.x * AL_WARPS + warp; g < total; g += gridDim.x * AL_WARPS) {
            int lo = 0, hi = n_active;              // slot j with pref[j] <= g < pref[j + 1]
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pref[mid] <= g) lo = mid; else hi = mid; }
            k5_tile<0>(p, s_slot[lo], g - s_pref[lo], it, terms[warp], lane);
        }
        grid.sync();
        for (int j = blockIdx.x; j < n_active; j += gridDim.x) k5_solve_slot(p, s_slot[j], it, M);
        grid.sync();
    }
}

__global__ void k5_init(MapParams p, const float *seeds, int n)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < 2 * p.max_iter + 1; i += blockDim.x) p.done[2 * n + i] = 0;   // converged count, tile tickets
    if (k >= n) return;
    for (int a = 0; a < 6; a++) p.T[k * 6 + a] = seeds[k * 6 + a];
    int *st = p.state + k * 4;
    st[0] = 0; st[1] = 0; st[2] = 0; st[3] = VLO_SOFT_TOO_FEW_CORR;
    p.ncorr[k * 2] = 0; p.ncorr[k * 2 + 1] = 0; p.done[2 * k] = 0; p.done[2 * k + 1] = K5_NOT_CONVERGED;
    vlo_result *r = p.result + k;
    for (int a = 0; a < 6; a++) { r->transform[a] = seeds[k * 6 + a]; r->eig[a] = 0.0f; }
    for (int a = 0; a < 36; a++) { r->hessian[a] = 0.0f; r->P[a] = (a % 7 == 0) ? 1.0f : 0.0f; r->cov[a] = 0.0; }
    r->is_degenerate = 0; r->iterations = 0; r->n_corr_edge = 0; r->n_corr_plane = 0;
    r->logdet_rot = 0.f; r->logdet_trans = 0.f; r->pass_dopt = 0; r->status = VLO_SOFT_TOO_FEW_CORR;
}

int vlo_launch_register_map(vlo_handle *h, const int *d_scans, int n, const float *d_seeds)
{
    ScanBatchDev &sb = h->sb; const vlo_config &c = h->cfg;
    MapParams p;
    // queries = the sweep's down-sampled corner / surface stacks (k7_map.cu; same strides as the feature clouds)
    (void)sb;
    p.lsharp_pts = h->lm.ds_pts[0]; p.cap_lsharp = h->cap_lsharp; p.lflat_pts = h->lm.ds_pts[1]; p.N = c.max_points;
    p.counts = h->lm.ds_counts; p.n_rings = c.n_rings;
    p.scans = d_scans; p.T = h->map_T; p.state = h->map_state; p.result = h->map_result;
    p.idx5 = h->map_idx5; p.qcap = h->cap_lsharp + c.max_points;
    p.partials = h->map_partials; p.pcap = (p.qcap + 31) / 32 + 8; p.ncorr = h->map_ncorr; p.done = h->map_done;
    p.gm0 = h->gs_map[0]; p.gm1 = h->gs_map[1]; p.rho0 = grid_thread_rho(p.gm0.cell, 1.0f); p.rho1 = grid_thread_rho(p.gm1.cell, 1.0f); p.map0 = h->map_pts[0]; p.map1 = h->map_pts[1]; p.map_n = h->lm.mode == 2 ? h->lm.sub_n : h->map_n;
    p.max_iter = c.map_max_iterations; p.degen_thr = c.map_degen_eig; p.dT_abort = c.map_delta_t_abort;
    p.dR_abort = c.map_delta_r_abort; p.rot_thr = c.dopt_rot_threshold; p.trans_thr = c.dopt_trans_threshold;
    k5_init<<<(n + 127) / 128, 128, 0, h->stream>>>(p, d_seeds, n);
    int qmax = h->map_qmax > 0 ? h->map_qmax : p.qcap;
    if (n <= VLO_COOP_MAX_SLOTS) {
        // latency path: one cooperative launch, as many co-resident CTAs as there can be warp tiles
        int &coresident = h->coop_resident;     // co-resident CTA capacity of k5_register_coop for this handle's shared-memory size
        const size_t dyn = sizeof(int) * (2 * (size_t)n + 2);
        if (!coresident) {
            int per_sm = 0, sms = 0;
            VLO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k5_register_coop, KNN_THREADS, sizeof(int) * (2 * (size_t)VLO_COOP_MAX_SLOTS + 2)));
            VLO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device));
            coresident = std::max(1, per_sm * sms);
        }
        long long tiles = (long long)n * ((qmax + 31) / 32);
        int ctas = (int)std::min<long long>(coresident, std::max<long long>(1, (tiles + AL_WARPS - 1) / AL_WARPS));
        void *args[] = { (void *)&p, (void *)&n };
        vlo_prof_begin(h, ST_MAP_KNN);
        VLO_CUDA(cudaLaunchCooperativeKernel((void *)k5_register_coop, dim3(ctas), dim3(KNN_THREADS), args, dyn, h->stream));
        vlo_prof_end(h, ST_MAP_KNN);
        h->launches += 2;
    } else {
        // throughput path: two launches per Gauss-Newton iteration (association, then linearisation + solve), each a
        // persistent grid over the flat tile list of the unconverged slots, never more warps than tiles
        int &sms = h->dev_sms, &occ_a = h->k5_occ_assoc, &occ_l = h->k5_occ_lin;      // per handle = per device
        if (!sms) {
            VLO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device));
            VLO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, k5_assoc, KNN_THREADS, 2048));
            VLO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_l, k5_lin, KNN_THREADS, 2048));
        }
        const long long tiles = (long long)n * ((qmax + 31) / 32);
        const long long want = (tiles + AL_WARPS - 1) / AL_WARPS;
        const int ctas_a = (int)std::max<long long>(1, std::min<long long>((long long)sms * std::max(1, occ_a), want));
        const int ctas_l = (int)std::max<long long>(1, std::min<long long>((long long)sms * std::max(1, occ_l), want));
        const size_t dyn = sizeof(int) * (2 * (size_t)n + 2);
        if (dyn > 24 * 1024) {          // the slot list lives in shared memory next to 22 KB of static scratch
            h->err = "vlo_register_map: at most 3000 scans per call (split the batch)";
            return VLO_ERR_CAPACITY;
        }
        for (int it = 0; it < c.map_max_iterations; it++) {
            VLO_PROF(h, ST_MAP_ASSOC, (k5_assoc<<<ctas_a, KNN_THREADS, dyn, h->stream>>>(p, it, n)));
            VLO_PROF(h, ST_MAP_LIN, (k5_lin<<<ctas_l, KNN_THREADS, dyn, h->stream>>>(p, it, n)));
        }
        h->launches += 1 + 2 * c.map_max_iterations;
    }
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}
