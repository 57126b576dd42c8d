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
#include <algorithm>
#include <cooperative_groups.h>

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

__device__ __forceinline__ float4 to_map(const float *T, const float *trig, float4 pi)
{
    float x = pi.x, y = pi.y, z = pi.z;
    float sx = trig[0], cx = trig[1], sy = trig[2], cy = trig[3], sz = trig[4], cz = trig[5];
    float x0 = x; x = cz * x0 - sz * y; y = sz * x0 + cz * y;
    float y0 = y; y = cx * y0 - sx * z; z = sx * y0 + cx * z;
    x0 = x;       x = cy * x0 + sy * z; z = cy * z - sy * x0;
    return make_float4(x + T[3], y + T[4], z + T[5], pi.w);
}

// cyclic Jacobi on a symmetric 3x3; eval ascending, evec[k*3+i] = component i of eigenvector k
__device__ inline void eig3_jacobi(const float *Ain, float *eval, float *evec)
{
    float A[3][3], V[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { A[i][j] = Ain[i * 3 + j]; V[i][j] = (i == j) ? 1.0f : 0.0f; }
    for (int sweep = 0; sweep < 6; sweep++) {
        #pragma unroll
        for (int m = 0; m < 3; m++) {
            const int p = (m == 2) ? 1 : 0, q = (m == 0) ? 1 : 2;
            float apq = A[p][q];
            if (apq == 0.0f) continue;
            float theta = (A[q][q] - A[p][p]) / (2.0f * apq);
            float t = 1.0f / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
            if (theta < 0.0f) t = -t;
            float c = 1.0f / sqrtf(t * t + 1.0f), s = t * c;
            #pragma unroll
            for (int i = 0; i < 3; i++) { float aip = A[i][p], aiq = A[i][q]; A[i][p] = c * aip - s * aiq; A[i][q] = s * aip + c * aiq; }
            #pragma unroll
            for (int j = 0; j < 3; j++) { float apj = A[p][j], aqj = A[q][j]; A[p][j] = c * apj - s * aqj; A[q][j] = s * apj + c * aqj; }
            #pragma unroll
            for (int i = 0; i < 3; i++) { float vip = V[i][p], viq = V[i][q]; V[i][p] = c * vip - s * viq; V[i][q] = s * vip + c * viq; }
            A[q][p] = A[p][q];
            const int r = 3 - p - q;
            A[r][p] = A[p][r]; A[r][q] = A[q][r];
        }
    }
    // stable insertion sort of the three eigenpairs by eigenvalue (strict <), written as adjacent conditional swaps on
    // named registers: no run-time array index, so nothing of this function lives in local memory
    float e0 = A[0][0], e1 = A[1][1], e2 = A[2][2];
    float c0[3] = { V[0][0], V[1][0], V[2][0] }, c1[3] = { V[0][1], V[1][1], V[2][1] }, c2[3] = { V[0][2], V[1][2], V[2][2] };
    #define EIG3_SWAP(ea, ca, eb, cb) { float t_ = ea; ea = eb; eb = t_; _Pragma("unroll") for (int i_ = 0; i_ < 3; i_++) { float u_ = ca[i_]; ca[i_] = cb[i_]; cb[i_] = u_; } }
    if (e1 < e0) EIG3_SWAP(e0, c0, e1, c1)
    if (e2 < e1) { EIG3_SWAP(e1, c1, e2, c2) if (e1 < e0) EIG3_SWAP(e0, c0, e1, c1) }
    #undef EIG3_SWAP
    eval[0] = e0; eval[1] = e1; eval[2] = e2;
    #pragma unroll
    for (int i = 0; i < 3; i++) { evec[i] = c0[i]; evec[3 + i] = c1[i]; evec[6 + i] = c2[i]; }
}

// One Householder step of the 5x3 column-pivoted QR with every index a compile-time constant (K = step).
template <int K>
__device__ __forceinline__ void lstsq53_step(float (&A)[5][3], float (&b)[5], int (&perm)[3], int &nonzero, float thr_helper)
{
    int piv = K; float best = -1.0f;
    #pragma unroll
    for (int j = K; j < 3; j++) {
        float s = 0.0f;
        #pragma unroll
        for (int i = K; i < 5; i++) s += A[i][j] * A[i][j];
        if (s > best) { best = s; piv = j; }
    }
    if (nonzero == 3 && best < thr_helper * (float)(5 - K)) nonzero = K;
    #pragma unroll
    for (int j = K + 1; j < 3; j++)
        if (piv == j) {
            #pragma unroll
            for (int i = 0; i < 5; i++) { float t = A[i][K]; A[i][K] = A[i][j]; A[i][j] = t; }
            int t = perm[K]; perm[K] = perm[j]; perm[j] = t;
        }
    const float nrm = sqrtf(best);
    if (nrm == 0.0f) return;
    const float alpha = (A[K][K] >= 0.0f) ? -nrm : nrm;
    float v[5];
    #pragma unroll
    for (int i = K; i < 5; i++) v[i] = A[i][K];
    v[K] = v[K] - alpha;
    float vn2 = 0.0f;
    #pragma unroll
    for (int i = K; i < 5; i++) vn2 += v[i] * v[i];
    if (vn2 == 0.0f) return;
    #pragma unroll
    for (int j = K; j < 3; j++) {
        float dot = 0.0f;
        #pragma unroll
        for (int i = K; i < 5; i++) dot += v[i] * A[i][j];
        const float f = (2.0f * dot) / vn2;
        #pragma unroll
        for (int i = K; i < 5; i++) A[i][j] = A[i][j] - f * v[i];
    }
    float dot = 0.0f;
    #pragma unroll
    for (int i = K; i < 5; i++) dot += v[i] * b[i];
    const float f = (2.0f * dot) / vn2;
    #pragma unroll
    for (int i = K; i < 5; i++) b[i] = b[i] - f * v[i];
}

// min ||A x + 1|| for the 5 x 3 neighbour matrix: column-pivoted Householder QR (Q1 of the oracle, orc_lstsq53), same
// operations in the same order; the pivot swaps, the rank-dependent back-substitution and the final permutation are
// spelled out on constant indices so that A, b, v stay in registers (the run-time-indexed version kept them in local
// memory: 623 LDL/STL in the kernel)
__device__ __forceinline__ void lstsq53(const float (*Ain)[3], float *x)
{
    float A[5][3], b[5];
    int perm[3] = { 0, 1, 2 };
    #pragma unroll
    for (int i = 0; i < 5; i++) {
        #pragma unroll
        for (int j = 0; j < 3; j++) A[i][j] = Ain[i][j];
        b[i] = -1.0f;
    }
    float maxn2 = 0.0f;
    #pragma unroll
    for (int j = 0; j < 3; j++) {
        float s = 0.0f;
        #pragma unroll
        for (int i = 0; i < 5; i++) s += A[i][j] * A[i][j];
        if (s > maxn2) maxn2 = s;
    }
    const float mx = sqrtf(maxn2) * FLT_EPSILON;
    const float thr_helper = (mx * mx) / 5.0f;
    int nonzero = 3;
    lstsq53_step<0>(A, b, perm, nonzero, thr_helper);
    lstsq53_step<1>(A, b, perm, nonzero, thr_helper);
    lstsq53_step<2>(A, b, perm, nonzero, thr_helper);
    float y0 = 0.f, y1 = 0.f, y2 = 0.f;
    if (nonzero == 3) {
        y2 = b[2] / A[2][2];
        y1 = (b[1] - A[1][2] * y2) / A[1][1];
        float s0 = b[0]; s0 = s0 - A[0][1] * y1; s0 = s0 - A[0][2] * y2;
        y0 = s0 / A[0][0];
    } else if (nonzero == 2) {
        y1 = b[1] / A[1][1];
        y0 = (b[0] - A[0][1] * y1) / A[0][0];
    } else if (nonzero == 1) {
        y0 = b[0] / A[0][0];
    }
    #pragma unroll
    for (int c = 0; c < 3; c++) x[c] = (perm[0] == c) ? y0 : ((perm[1] == c) ? y1 : y2);
}

__device__ inline bool map_edge_coeff(float4 sel, const float4 *nb, float *coeff)
{
    float vx = 0.0f, vy = 0.0f, vz = 0.0f;
    for (int j = 0; j < 5; j++) { vx += nb[j].x; vy += nb[j].y; vz += nb[j].z; }
    vx = vx / 5.0f; vy = vy / 5.0f; vz = vz / 5.0f;
    float a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
    for (int j = 0; j < 5; j++) {
        float ax = nb[j].x - vx, ay = nb[j].y - vy, az = nb[j].z - vz;
        a00 += ax * ax; a10 += ax * ay; a20 += ax * az; a11 += ay * ay; a21 += ay * az; a22 += az * az;
    }
    float M[9];
    M[0] = a00 / 5.0f; M[4] = a11 / 5.0f; M[8] = a22 / 5.0f;
    M[3] = M[1] = a10 / 5.0f; M[6] = M[2] = a20 / 5.0f; M[7] = M[5] = a21 / 5.0f;
    float ev[3], evec[9];
    eig3_jacobi(M, ev, evec);
    if (!(ev[2] > 3.0f * ev[1])) return false;
    float x0 = sel.x, y0 = sel.y, z0 = sel.z;
    float x1 = (float)((double)vx + 0.1 * (double)evec[6]), y1 = (float)((double)vy + 0.1 * (double)evec[7]), z1 = (float)((double)vz + 0.1 * (double)evec[8]);
    float x2 = (float)((double)vx - 0.1 * (double)evec[6]), y2 = (float)((double)vy - 0.1 * (double)evec[7]), z2 = (float)((double)vz - 0.1 * (double)evec[8]);
    float m1 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    float m2 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    float m3 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    float a012 = sqrtf(m1 * m1 + m2 * m2 + m3 * m3);
    float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    float la = ((y1 - y2) * m1 + (z1 - z2) * m2) / a012 / l12;
    float lb = -((x1 - x2) * m1 - (z1 - z2) * m3) / a012 / l12;
    float lc = -((x1 - x2) * m2 + (y1 - y2) * m3) / a012 / l12;
    float ld2 = a012 / l12;
    float s = 1.0f - 0.9f * fabsf(ld2);
    coeff[0] = s * la; coeff[1] = s * lb; coeff[2] = s * lc; coeff[3] = s * ld2;
    return (double)s > 0.1;
}

__device__ inline bool map_plane_coeff(float4 sel, const float4 *nb, float *coeff)
{
    float A0[5][3], X0[3];
    for (int j = 0; j < 5; j++) { A0[j][0] = nb[j].x; A0[j][1] = nb[j].y; A0[j][2] = nb[j].z; }
    lstsq53(A0, X0);
    float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1.0f;
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa = pa / ps; pb = pb / ps; pc = pc / ps; pd = pd / ps;
    for (int j = 0; j < 5; j++)
        if ((double)fabsf(pa * nb[j].x + pb * nb[j].y + pc * nb[j].z + pd) > 0.2) return false;
    float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
    float s = 1.0f - 0.9f * fabsf(pd2) / sqrtf(sqrtf(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z));
    coeff[0] = s * pa; coeff[1] = s * pb; coeff[2] = s * pc; coeff[3] = s * pd2;
    return (double)s > 0.1;
}

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
__global__ void __launch_bounds__(KNN_THREADS) k5_assoc(MapParams p, int it, int n)
{
    extern __shared__ int s_dyn[];                 // [n] active slots, [n + 1] tile prefix
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
__global__ void __launch_bounds__(KNN_THREADS, 6) k5_lin(MapParams p, int it, int n)
{
    extern __shared__ int s_dyn[];
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
    extern __shared__ int s_dyn[];                 // [n] active slots, [n + 1] tile prefix
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
        static int sms = 0, occ_a = 0, occ_l = 0;
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
