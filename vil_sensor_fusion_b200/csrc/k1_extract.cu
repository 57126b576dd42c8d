// K1  feature extraction: one CTA per (ring, scan).  Replaces BasicScanRegistration::extractFeatures
// of the `loam` nodelet (gtsam_fusion/launch/loam.launch:33-38; knobs loam_params.yaml:25-31);
// SURVEY.md Appendix A.2-A.3 is the algorithm.  The ring is staged once in shared memory
// (coalesced float4 loads -> SoA) and everything else happens on chip:
//   B  occlusion / parallel-beam flags as a gather (window-OR) instead of upstream's scatter
//   C  11-tap curvature in upstream's summation order
//   D  greedy sector selection: no sort -- each pick is a warp arg-max / arg-min over the sector
//      ((curvature, index) lexicographic == walking upstream's stable insertion sort), REDUX for
//      the reduction, ballot for the +-K neighbour suppression.  Sectors run concurrently on
//      separate warps; the only cross-sector dependency (suppression marks spilling into the next
//      sector's first K points) is checked afterwards and the rare conflicting sector is redone.
//   F  less-flat voxel filter (k1c_lessflat, a kernel of its own so that neither phase's shared memory limits the
//      other's occupancy): shared-memory hash on the voxel triple, integer (order-free) sums, points kept in
//      registers, output in order of first appearance (ballot ranks + one 64-entry scan).
#include "vlo_internal.cuh"
#include <algorithm>

#define K1_THREADS 256
#define K1_EMPTY 0xFFFFFFFFFFFFFFFFull

struct K1Params {
    const float4 *cloud; const int *ring_start; int N; int n_rings;
    int K, NR, max_sharp, max_lsharp, max_flat; float thr; float leaf;
    int MR, HT; int scan_first;
    int8_t *label; float *curvature; uint8_t *picked;
    int *slot_sharp, *slot_lsharp, *slot_flat; uint8_t *slot_cnt;
    float4 *lflat_slotted; int *lflat_cnt;
    int *status_word;
};

struct K1Smem {
    float *x, *y, *z;
    float *curv;
    uint8_t *flag;          // bit2 f3 (until B2), bit3 gap(i,i+1) > 0.05, bit4 base picked
    uint8_t *mark0, *mark1; // suppression marks written by even / odd sectors
    int8_t *label;
    unsigned *w1, *w2;      // occlusion flags f1 / f2 as one bit per point (ballot words)
};

__device__ __forceinline__ K1Smem k1_carve(unsigned char *base, int MR)
{
    K1Smem s;
    s.x = (float *)base; base += (size_t)MR * 4;
    s.y = (float *)base; base += (size_t)MR * 4;
    s.z = (float *)base; base += (size_t)MR * 4;
    s.curv = (float *)base; base += (size_t)MR * 4;
    s.w1 = (unsigned *)base; base += (size_t)(MR / 32 + 2) * 4;
    s.w2 = (unsigned *)base; base += (size_t)(MR / 32 + 2) * 4;
    s.flag = base; base += MR;
    s.mark0 = base; base += MR;
    s.mark1 = base; base += MR;
    s.label = (int8_t *)base;
    return s;
}

static size_t k1_smem_bytes(int MR) { return (size_t)MR * 16 + (size_t)(MR / 32 + 2) * 8 + (size_t)MR * 4; }

// less-flat voxel hash (k1c_lessflat)
struct K1cSmem {
    unsigned long long *vkey;
    int *vfirst, *vcnt, *vsx, *vsy, *vsz, *vsw;
};

__device__ __forceinline__ K1cSmem k1c_carve(unsigned char *base, int HT)
{
    K1cSmem s;
    s.vkey = (unsigned long long *)base; base += (size_t)HT * 8;
    s.vfirst = (int *)base; base += (size_t)HT * 4;
    s.vcnt = (int *)base; base += (size_t)HT * 4;
    s.vsx = (int *)base; base += (size_t)HT * 4;
    s.vsy = (int *)base; base += (size_t)HT * 4;
    s.vsz = (int *)base; base += (size_t)HT * 4;
    s.vsw = (int *)base;
    return s;
}

static size_t k1c_smem_bytes(int HT) { return (size_t)HT * 32; }

// one sector's greedy selection, executed by one full warp.
// view_prev_lo..view_prev_hi: index range (ring-relative) where the previous sector's marks are visible.
__device__ void k1_sector_greedy(const K1Params &p, const K1Smem &s, int lane, int sec, int sp, int ep, int start,
                                 uint8_t *own, const uint8_t *prev, int view_prev_lo, int view_prev_hi, int b, int r)
{
    // sp, ep: ring-relative inclusive sector range; own: mark array this sector writes and sees
    const int K = p.K;
    size_t slot_base = ((size_t)(b * p.n_rings + r) * p.NR + sec);
    int n_sharp = 0, n_ls = 0, n_flat = 0;
    // ---- corners: largest (curvature, index) first
    for (int pick = 0; pick < p.max_lsharp; pick++) {
        unsigned best_c = 0u; int best_i = -1;
        for (int i = sp + lane; i <= ep; i += 32) {
            bool pk = (s.flag[i] & 16) || own[i] || (i >= view_prev_lo && i <= view_prev_hi && prev[i]);
            float c = s.curv[i];
            if (!pk && c > p.thr) {
                unsigned cb = __float_as_uint(c);
                if (cb > best_c || (cb == best_c && i > best_i)) { best_c = cb; best_i = i; }
            }
        }
        unsigned m = __reduce_max_sync(0xffffffffu, best_c);
        if (m == 0u) break;
        int cand = (best_c == m) ? best_i : -1;
        int idx = __reduce_max_sync(0xffffffffu, cand);
        // label + slot
        if (lane == 0) {
            if (pick < p.max_sharp) { s.label[idx] = 2; p.slot_sharp[slot_base * p.max_sharp + n_sharp] = start + idx; }
            else s.label[idx] = 1;
            p.slot_lsharp[slot_base * p.max_lsharp + n_ls] = start + idx;
        }
        if (pick < p.max_sharp) n_sharp++;
        n_ls++;
        // markAsPicked
        unsigned g = 1u;
        if (lane < K) g = (s.flag[idx + lane] >> 3) & 1u;                 // gap(idx+lane, idx+lane+1)
        else if (lane >= 16 && lane < 16 + K) g = (s.flag[idx - 1 - (lane - 16)] >> 3) & 1u;
        unsigned bal = __ballot_sync(0xffffffffu, g);
        int fcount = min(K, __ffs(bal & 0xffffu) - 1);
        int bcount = min(K, __ffs(bal >> 16) - 1);
        if (lane == 0) own[idx] = 1;
        if (lane < fcount) own[idx + 1 + lane] = 1;
        if (lane >= 16 && lane - 16 < bcount) own[idx - 1 - (lane - 16)] = 1;
        __syncwarp();
    }
    // ---- flats: smallest (curvature, index) first
    for (int pick = 0; pick < p.max_flat; pick++) {
        unsigned best_c = 0xffffffffu; int best_i = 0x7fffffff;
        for (int i = sp + lane; i <= ep; i += 32) {
            bool pk = (s.flag[i] & 16) || own[i] || (i >= view_prev_lo && i <= view_prev_hi && prev[i]);
            float c = s.curv[i];
            if (!pk && c < p.thr) {
                unsigned cb = __float_as_uint(c);
                if (cb < best_c || (cb == best_c && i < best_i)) { best_c = cb; best_i = i; }
            }
        }
        unsigned m = __reduce_min_sync(0xffffffffu, best_c);
        if (m == 0xffffffffu) break;
        int cand = (best_c == m) ? best_i : 0x7fffffff;
        int idx = __reduce_min_sync(0xffffffffu, cand);
        if (lane == 0) { s.label[idx] = -1; p.slot_flat[slot_base * p.max_flat + n_flat] = start + idx; }
        n_flat++;
        unsigned g = 1u;
        if (lane < K) g = (s.flag[idx + lane] >> 3) & 1u;
        else if (lane >= 16 && lane < 16 + K) g = (s.flag[idx - 1 - (lane - 16)] >> 3) & 1u;
        unsigned bal = __ballot_sync(0xffffffffu, g);
        int fcount = min(K, __ffs(bal & 0xffffu) - 1);
        int bcount = min(K, __ffs(bal >> 16) - 1);
        if (lane == 0) own[idx] = 1;
        if (lane < fcount) own[idx + 1 + lane] = 1;
        if (lane >= 16 && lane - 16 < bcount) own[idx - 1 - (lane - 16)] = 1;
        __syncwarp();
    }
    if (lane == 0) {
        uint8_t *c = p.slot_cnt + slot_base * 4;
        c[0] = (uint8_t)n_sharp; c[1] = (uint8_t)n_ls; c[2] = (uint8_t)n_flat; c[3] = 0;
    }
}

// Register-resident variant (sectors of up to 32*MAXS points, i.e. every real lidar): each lane keeps
// the curvature of its lane-strided points and two eligibility bit masks in registers, so a pick is
// a local scan over MAXS registers + 2 REDUX + 1 ballot; shared memory is touched only to publish
// the suppression marks.  Same selection order as k1_sector_greedy, bit for bit.
template <int MAXS>
__device__ void k1_sector_greedy_regs(const K1Params &p, const K1Smem &s, int lane, int sec, int sp, int ep, int start,
                                      uint8_t *own, const uint8_t *prev, int view_prev_lo, int view_prev_hi, int b, int r)
{
    const int K = p.K;
    size_t slot_base = ((size_t)(b * p.n_rings + r) * p.NR + sec);
    unsigned cb[MAXS];
    unsigned el_c = 0u, el_f = 0u;            // eligibility: corner (c > thr) / flat (c < thr), not picked
    #pragma unroll
    for (int k = 0; k < MAXS; k++) {
        int i = sp + lane + 32 * k;
        cb[k] = 0u;
        if (i <= ep) {
            float c = s.curv[i];
            cb[k] = __float_as_uint(c);
            bool pk = (s.flag[i] & 16) || own[i] || (i >= view_prev_lo && i <= view_prev_hi && prev[i]);
            if (!pk && c > p.thr) el_c |= 1u << k;
            if (!pk && c < p.thr) el_f |= 1u << k;
        }
    }
    int n_sharp = 0, n_ls = 0, n_flat = 0;
    for (int phase = 0; phase < 2; phase++) {
        const int max_picks = phase == 0 ? p.max_lsharp : p.max_flat;
        for (int pick = 0; pick < max_picks; pick++) {
            int idx;
            if (phase == 0) {                 // largest (curvature, index)
                unsigned best_c = 0u; int best_k = -1;
                #pragma unroll
                for (int k = 0; k < MAXS; k++) if (((el_c >> k) & 1u) && cb[k] >= best_c) { best_c = cb[k]; best_k = k; }
                unsigned m = __reduce_max_sync(0xffffffffu, best_k >= 0 ? best_c : 0u);
                if (m == 0u) break;
                int cand = (best_k >= 0 && best_c == m) ? sp + lane + 32 * best_k : -1;
                idx = __reduce_max_sync(0xffffffffu, cand);
            } else {                          // smallest (curvature, index)
                unsigned best_c = 0xffffffffu; int best_k = -1;
                #pragma unroll
                for (int k = MAXS - 1; k >= 0; k--) if (((el_f >> k) & 1u) && cb[k] <= best_c) { best_c = cb[k]; best_k = k; }
                unsigned m = __reduce_min_sync(0xffffffffu, best_k >= 0 ? best_c : 0xffffffffu);
                if (m == 0xffffffffu) break;
                int cand = (best_k >= 0 && best_c == m) ? sp + lane + 32 * best_k : 0x7fffffff;
                idx = __reduce_min_sync(0xffffffffu, cand);
            }
            if (lane == 0) {
                if (phase == 0) {
                    if (pick < p.max_sharp) { s.label[idx] = 2; p.slot_sharp[slot_base * p.max_sharp + n_sharp] = start + idx; }
                    else s.label[idx] = 1;
                    p.slot_lsharp[slot_base * p.max_lsharp + n_ls] = start + idx;
                } else {
                    s.label[idx] = -1; p.slot_flat[slot_base * p.max_flat + n_flat] = start + idx;
                }
            }
            if (phase == 0) { if (pick < p.max_sharp) n_sharp++; n_ls++; } else n_flat++;
            // markAsPicked: +-K neighbours while consecutive gaps stay <= 0.05
            unsigned g = 1u;
            if (lane < K) g = (s.flag[idx + lane] >> 3) & 1u;
            else if (lane >= 16 && lane < 16 + K) g = (s.flag[idx - 1 - (lane - 16)] >> 3) & 1u;
            unsigned bal = __ballot_sync(0xffffffffu, g);
            int fcount = min(K, __ffs(bal & 0xffffu) - 1);
            int bcount = min(K, __ffs(bal >> 16) - 1);
            if (lane == 0) own[idx] = 1;
            if (lane < fcount) own[idx + 1 + lane] = 1;
            if (lane >= 16 && lane - 16 < bcount) own[idx - 1 - (lane - 16)] = 1;
            // clear eligibility of my point inside [idx - bcount, idx + fcount] (at most one per lane: width <= 2K+1 < 32)
            int lo = idx - bcount, hi = idx + fcount;
            int k0 = (lo - sp - lane + 31) >> 5;            // first k with sp + lane + 32k >= lo
            if (lo - sp - lane < 0) k0 = 0;
            int i0 = sp + lane + 32 * k0;
            if (k0 < MAXS && i0 >= lo && i0 <= hi) { el_c &= ~(1u << k0); el_f &= ~(1u << k0); }
        }
    }
    __syncwarp();
    if (lane == 0) {
        uint8_t *c = p.slot_cnt + slot_base * 4;
        c[0] = (uint8_t)n_sharp; c[1] = (uint8_t)n_ls; c[2] = (uint8_t)n_flat; c[3] = 0;
    }
}

#define K1_MAXS 12
__device__ __forceinline__ void k1_greedy_dispatch(const K1Params &p, const K1Smem &s, int lane, int sec, int sp, int ep, int start,
                                                   uint8_t *own, const uint8_t *prev, int vlo, int vhi, int b, int r)
{
    if (ep - sp + 1 <= 32 * K1_MAXS) k1_sector_greedy_regs<K1_MAXS>(p, s, lane, sec, sp, ep, start, own, prev, vlo, vhi, b, r);
    else k1_sector_greedy(p, s, lane, sec, sp, ep, start, own, prev, vlo, vhi, b, r);
}

#ifdef VLO_HOST_EMULATION                      // tests/host/: the CTA's dynamic shared memory is a plain array
alignas(16) static unsigned char k1_smem_raw[256 * 1024];
#else
extern __shared__ __align__(16) unsigned char k1_smem_raw[];
#endif

// sector bounds of a ring (upstream's integer arithmetic on absolute indices); returns whether every sector is non-empty
__device__ __forceinline__ void k1_sector_bounds(int start, int n, int K, int NR, int tid, int *s_sp, int *s_ep, int *s_all_valid)
{
    if (tid < NR) {
        int a = start + K, e = start + n - 1 - K;
        int sp = (a * (NR - tid) + e * tid) / NR;
        int ep = (a * (NR - 1 - tid) + e * (tid + 1)) / NR - 1;
        s_sp[tid] = sp - start; s_ep[tid] = ep - start;
    }
    if (tid == 0) {
        int a = start + K, e = start + n - 1 - K, ok = 1;
        for (int j = 0; j < NR; j++) {
            int sp = (a * (NR - j) + e * j) / NR, ep = (a * (NR - 1 - j) + e * (j + 1)) / NR - 1;
            if (!(ep > sp)) ok = 0;
        }
        *s_all_valid = ok;        // every sector non-empty: "inside some sector" == inside [sp_0, ep_last]
    }
}

__global__ void __launch_bounds__(K1_THREADS) k1_extract(K1Params p)
{
    __shared__ int s_sp[VLO_MAX_REGIONS], s_ep[VLO_MAX_REGIONS];
    __shared__ int s_seq, s_all_valid;
    const int r = blockIdx.x, b = p.scan_first + blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int start = p.ring_start[b * (VLO_MAX_RINGS + 1) + r];
    const int n = p.ring_start[b * (VLO_MAX_RINGS + 1) + r + 1] - start;
    const int K = p.K, NR = p.NR;
    const size_t gbase = (size_t)b * p.N + start;
    K1Smem s = k1_carve(k1_smem_raw, p.MR);

    // slot counters default to zero
    if (tid < NR * 4) p.slot_cnt[((size_t)(b * p.n_rings + r) * NR) * 4 + tid] = 0;
    if (tid == 0) p.lflat_cnt[b * p.n_rings + r] = 0;
    if (n <= 0) return;
    bool skip = (n <= 2 * K + 1) || (n > p.MR);
    if (n > p.MR && tid == 0) atomicOr(p.status_word, 1);
    if (skip) {
        for (int i = tid; i < n; i += K1_THREADS) { p.label[gbase + i] = 0; p.curvature[gbase + i] = 0.f; p.picked[gbase + i] = 0; }
        return;
    }
    // ---- A: stage ring
    for (int i = tid; i < n; i += K1_THREADS) {
        float4 v = p.cloud[gbase + i];
        s.x[i] = v.x; s.y[i] = v.y; s.z[i] = v.z;
        s.label[i] = 0; s.mark0[i] = 0; s.mark1[i] = 0;
    }
    for (int k = tid; k < p.MR / 32 + 2; k += K1_THREADS) { s.w1[k] = 0u; s.w2[k] = 0u; }
    if (tid == 0) s_seq = 0;
    k1_sector_bounds(start, n, K, NR, tid, s_sp, s_ep, &s_all_valid);
    __syncthreads();
    // ---- B1: per-point flags; f1 / f2 (occlusion marks that spread over K neighbours) go to one-bit-per-point words
    for (int base = warp * 32; base < n; base += K1_THREADS) {
        const int i = base + lane;
        unsigned f = 0;
        if (i < n) {
            float px = s.x[i], py = s.y[i], pz = s.z[i];
            float diffNext = 0.f;
            if (i + 1 < n) {
                diffNext = sqdiff3(s.x[i + 1], s.y[i + 1], s.z[i + 1], px, py, pz);
                if ((double)diffNext > 0.05) f |= 8u;
            }
            if (i >= K && i < n - 1 - K) {
                bool cont = false;
                if ((double)diffNext > 0.1) {
                    float nx = s.x[i + 1], ny = s.y[i + 1], nz = s.z[i + 1];
                    float depth1 = sqrtf((px * px + py * py) + pz * pz);
                    float depth2 = sqrtf((nx * nx + ny * ny) + nz * nz);
                    if (depth1 > depth2) {
                        float wq = depth2 / depth1;
                        float wd = sqrtf(sqdiff3(nx, ny, nz, px * wq, py * wq, pz * wq)) / depth2;
                        if ((double)wd < 0.1) { f |= 1u; cont = true; }
                    } else {
                        float wq = depth1 / depth2;
                        float wd = sqrtf(sqdiff3(px, py, pz, nx * wq, ny * wq, nz * wq)) / depth1;
                        if ((double)wd < 0.1) f |= 2u;
                    }
                }
                if (!cont) {
                    float diffPrev = sqdiff3(px, py, pz, s.x[i - 1], s.y[i - 1], s.z[i - 1]);
                    float dis = (px * px + py * py) + pz * pz;
                    if ((double)diffNext > 0.0002 * (double)dis && (double)diffPrev > 0.0002 * (double)dis) f |= 4u;
                }
            }
            s.flag[i] = (uint8_t)(f & 12u);
        }
        const unsigned b1 = __ballot_sync(0xffffffffu, f & 1u), b2 = __ballot_sync(0xffffffffu, f & 2u);
        if (lane == 0) { s.w1[base >> 5] = b1; s.w2[base >> 5] = b2; }
    }
    __syncthreads();
    // ---- B2: window-OR -> base picked (bit4): f3(i) | any f1 in [i, i+K] | any f2 in [i-1-K, i-1];  C: curvature
    const int lo_all = s_sp[0], hi_all = s_ep[NR - 1];
    const bool all_valid = s_all_valid != 0;
    for (int i = tid; i < n; i += K1_THREADS) {
        const unsigned fl = s.flag[i];
        unsigned pk = (fl >> 2) & 1u;
        {
            const int j = i >> 5;
            const unsigned v = __funnelshift_r(s.w1[j], s.w1[j + 1], i & 31);
            pk |= (v & ((2u << K) - 1u)) != 0u;
        }
        {
            const int lo = max(i - 1 - K, 0), len = i - lo;
            if (len > 0) {
                const int j = lo >> 5;
                const unsigned v = __funnelshift_r(s.w2[j], s.w2[j + 1], lo & 31);
                pk |= (v & ((1u << len) - 1u)) != 0u;
            }
        }
        float cv = 0.f;
        if (i >= lo_all && i <= hi_all) {
            bool in = all_valid;
            if (!all_valid) for (int j = 0; j < NR; j++) in |= (s_ep[j] > s_sp[j] && i >= s_sp[j] && i <= s_ep[j]);
            if (in) {
                float wgt = (float)(-2 * K);
                float dx = wgt * s.x[i], dy = wgt * s.y[i], dz = wgt * s.z[i];
                for (int m = 1; m <= K; m++) {
                    dx += s.x[i + m] + s.x[i - m];
                    dy += s.y[i + m] + s.y[i - m];
                    dz += s.z[i + m] + s.z[i - m];
                }
                cv = (dx * dx + dy * dy) + dz * dz;
            }
        }
        s.curv[i] = cv;
        p.curvature[gbase + i] = cv;
        s.flag[i] = (uint8_t)((fl & 8u) | (pk << 4));       // nobody else reads this byte before the barrier
    }
    // decide concurrent vs sequential sector processing
    if (tid == 0) {
        int seq = 0;
        for (int j = 0; j < NR; j++) if (s_ep[j] > s_sp[j] && s_ep[j] - s_sp[j] + 1 < 2 * K + 2) seq = 1;
        if (NR > K1_THREADS / 32 || NR == 1) seq = 1;
        s_seq = seq;
    }
    __syncthreads();
    // ---- D: greedy selection
    if (!s_seq) {
        if (warp < NR && s_ep[warp] > s_sp[warp])
            k1_greedy_dispatch(p, s, lane, warp, s_sp[warp], s_ep[warp], start,
                               (warp & 1) ? s.mark1 : s.mark0, (warp & 1) ? s.mark0 : s.mark1, 1, 0, b, r);
        __syncthreads();
        if (warp == 0) {
            for (int j = 1; j < NR; j++) {
                if (!(s_ep[j] > s_sp[j])) continue;
                const uint8_t *prev = (j & 1) ? s.mark0 : s.mark1;
                uint8_t *own = (j & 1) ? s.mark1 : s.mark0;
                int i = s_sp[j] + lane;
                bool conflict = (lane < K) && (i <= s_ep[j]) && prev[i] && s.label[i] != 0;
                if (__any_sync(0xffffffffu, conflict)) {
                    for (int q = s_sp[j] - K + lane; q <= s_ep[j] + K; q += 32) if (q >= 0 && q < n) own[q] = 0;
                    for (int q = s_sp[j] + lane; q <= s_ep[j]; q += 32) s.label[q] = 0;
                    __syncwarp();
                    k1_greedy_dispatch(p, s, lane, j, s_sp[j], s_ep[j], start, own, prev, s_sp[j], s_sp[j] + K - 1, b, r);
                    __syncwarp();
                }
            }
        }
    } else if (warp == 0) {
        // short rings: sectors strictly in order on one warp, a single mark array sees everything
        for (int j = 0; j < NR; j++) {
            if (!(s_ep[j] > s_sp[j])) continue;
            k1_greedy_dispatch(p, s, lane, j, s_sp[j], s_ep[j], start, s.mark0, s.mark1, 1, 0, b, r);
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- E: labels + picked mask out
    for (int i = tid; i < n; i += K1_THREADS) {
        p.label[gbase + i] = s.label[i];
        p.picked[gbase + i] = (uint8_t)(((s.flag[i] >> 4) & 1u) | s.mark0[i] | s.mark1[i]);
    }
}

// ---- F: less-flat voxel filter of one ring (pcl::VoxelGrid, leaf lessFlatFilterSize; V1 / V2 of the oracle).
// Thread t keeps points t, t + 256, .. in registers (MAXP of them); the shared memory holds nothing but the hash,
// so three CTAs fit an SM.  Output order = first appearance of the voxel along the ring: leaders are ranked by
// (chunk of 256, warp, lane) with ballots and one scan of the MAXP x 8 warp counts.
template <int MAXP>
__global__ void __launch_bounds__(K1_THREADS) k1c_lessflat(K1Params p)
{
    __shared__ int s_sp[VLO_MAX_REGIONS], s_ep[VLO_MAX_REGIONS];
    __shared__ int s_all_valid;
    __shared__ int s_cnt[MAXP * (K1_THREADS / 32) + 1];
    const int r = blockIdx.x, b = p.scan_first + blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int start = p.ring_start[b * (VLO_MAX_RINGS + 1) + r];
    const int n = p.ring_start[b * (VLO_MAX_RINGS + 1) + r + 1] - start;
    const int K = p.K, NR = p.NR;
    if (n <= 2 * K + 1 || n > p.MR) return;              // k1_extract left lflat_cnt = 0 for these rings
    const size_t gbase = (size_t)b * p.N + start;
    K1cSmem s = k1c_carve(k1_smem_raw, p.HT);
    float4 v[MAXP]; int so[MAXP];
    // all of the thread's points and labels are requested up front (independent loads in flight together); they land
    // while the hash table is being cleared, and the insertions below run without waiting on global memory
    int8_t lab[MAXP];
    #pragma unroll
    for (int k = 0; k < MAXP; k++) {
        const int i = k * K1_THREADS + tid;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f); lab[k] = 1;
        if (i < n) { v[k] = p.cloud[gbase + i]; lab[k] = p.label[gbase + i]; }
    }
    for (int k = tid; k < p.HT; k += K1_THREADS) {
        s.vkey[k] = K1_EMPTY; s.vfirst[k] = 0x7fffffff; s.vcnt[k] = 0; s.vsx[k] = 0; s.vsy[k] = 0; s.vsz[k] = 0; s.vsw[k] = 0;
    }
    k1_sector_bounds(start, n, K, NR, tid, s_sp, s_ep, &s_all_valid);
    __syncthreads();
    const int lo_all = s_sp[0], hi_all = s_ep[NR - 1];
    const bool all_valid = s_all_valid != 0;
    const float leaf = p.leaf, inv = 1.0f / p.leaf;
    #pragma unroll
    for (int k = 0; k < MAXP; k++) {
        const int i = k * K1_THREADS + tid;
        so[k] = -1;
        bool in = false;
        if (i < n && i >= lo_all && i <= hi_all) {
            in = all_valid;
            if (!all_valid) for (int j = 0; j < NR; j++) in |= (s_ep[j] > s_sp[j] && i >= s_sp[j] && i <= s_ep[j]);
        }
        const bool take = in && lab[k] <= 0;
        // The lanes of a warp hold consecutive points of the ring, and consecutive points share a 0.2 m voxel more often
        // than not (3.5 cm apart at 10 m): a voxel shows up as a RUN of lanes.  Run heads come from one ballot, the
        // run's integer sums from a five-step segmented shuffle reduction, and the head lane alone touches the hash
        // (r01c ncu: 57 % of this kernel's shared-memory wavefronts were same-address atomic conflicts, the CAS probe
        // its top stall).  A voxel re-entered later is a second run with its own atomics; integer sums are order-free,
        // so the centroids stay bit-identical.
        const float4 q = v[k];
        int ix = 0, iy = 0, iz = 0, dx = 0, dy = 0, dz = 0, dw = 0;
        unsigned long long key = 0x8000000000000000ull | (unsigned long long)lane;      // private key: not taken
        if (take) {
            ix = (int)floorf(q.x * inv); iy = (int)floorf(q.y * inv); iz = (int)floorf(q.z * inv);
            key = ((unsigned long long)(unsigned)(ix + (1 << 20)) << 42) | ((unsigned long long)(unsigned)(iy + (1 << 20)) << 21)
                | (unsigned long long)(unsigned)(iz + (1 << 20));
            const float ox = (float)ix * leaf, oy = (float)iy * leaf, oz = (float)iz * leaf;
            dx = (int)rintf((q.x - ox) * 1048576.0f); dy = (int)rintf((q.y - oy) * 1048576.0f);
            dz = (int)rintf((q.z - oz) * 1048576.0f); dw = (int)rintf((q.w - (float)(int)q.w) * 1048576.0f);
        }
        const unsigned long long prevk = __shfl_up_sync(0xffffffffu, key, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prevk != key);
        const unsigned later = heads & ~((2u << lane) - 1u);
        const int run_end = later ? (__ffs(later) - 2) : 31;
        const int leader = 31 - __clz(heads & ((2u << lane) - 1u));
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int ax = __shfl_down_sync(0xffffffffu, dx, d), ay = __shfl_down_sync(0xffffffffu, dy, d);
            const int az = __shfl_down_sync(0xffffffffu, dz, d), aw = __shfl_down_sync(0xffffffffu, dw, d);
            if (lane + d <= run_end) { dx += ax; dy += ay; dz += az; dw += aw; }
        }
        int slot = -1;
        if (take && lane == leader) {
            const unsigned hsh = ((unsigned)ix * 73856093u) ^ ((unsigned)iy * 19349663u) ^ ((unsigned)iz * 83492791u);
            slot = (int)(hsh & (unsigned)(p.HT - 1));
            while (true) {
                unsigned long long old = atomicCAS(&s.vkey[slot], K1_EMPTY, key);
                if (old == K1_EMPTY || old == key) break;
                slot = (slot + 1) & (p.HT - 1);
            }
            atomicMin(&s.vfirst[slot], i);
            atomicAdd(&s.vcnt[slot], run_end - lane + 1);
            atomicAdd(&s.vsx[slot], dx);
            atomicAdd(&s.vsy[slot], dy);
            atomicAdd(&s.vsz[slot], dz);
            atomicAdd(&s.vsw[slot], dw);
        }
        slot = __shfl_sync(0xffffffffu, slot, leader);
        if (take) so[k] = slot;
    }
    __syncthreads();
    unsigned bal[MAXP];
    #pragma unroll
    for (int k = 0; k < MAXP; k++) {
        const bool lead = so[k] >= 0 && s.vfirst[so[k]] == k * K1_THREADS + tid;
        bal[k] = __ballot_sync(0xffffffffu, lead);
        if (lane == 0) s_cnt[k * (K1_THREADS / 32) + warp] = __popc(bal[k]);
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan of the MAXP * 8 counts in (chunk, warp) order
        constexpr int NE = MAXP * (K1_THREADS / 32), PER = (NE + 31) / 32;
        int loc[PER], sum = 0;
        #pragma unroll
        for (int e = 0; e < PER; e++) { const int idx = lane * PER + e; loc[e] = idx < NE ? s_cnt[idx] : 0; sum += loc[e]; }
        int inc = sum;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
        int run = inc - sum;
        #pragma unroll
        for (int e = 0; e < PER; e++) { const int idx = lane * PER + e; if (idx < NE) s_cnt[idx] = run; run += loc[e]; }
        if (lane == 31) s_cnt[NE] = inc;
    }
    __syncthreads();
    #pragma unroll
    for (int k = 0; k < MAXP; k++) {
        if (!((bal[k] >> lane) & 1u)) continue;
        const int slot = so[k];
        const int pos = s_cnt[k * (K1_THREADS / 32) + warp] + __popc(bal[k] & ((1u << lane) - 1u));
        const float4 q = v[k];
        const float c = (float)s.vcnt[slot];
        const int ix = (int)floorf(q.x * inv), iy = (int)floorf(q.y * inv), iz = (int)floorf(q.z * inv);
        const float ox = (float)ix * leaf, oy = (float)iy * leaf, oz = (float)iz * leaf;
        const float qq = 1.0f / 1048576.0f;
        float4 o;
        o.x = ox + ((float)s.vsx[slot] / c) * qq;
        o.y = oy + ((float)s.vsy[slot] / c) * qq;
        o.z = oz + ((float)s.vsz[slot] / c) * qq;
        o.w = (float)(int)q.w + ((float)s.vsw[slot] / c) * qq;
        p.lflat_slotted[gbase + pos] = o;
    }
    if (tid == 0) p.lflat_cnt[b * p.n_rings + r] = s_cnt[MAXP * (K1_THREADS / 32)];
}

// K1b  compaction: per scan, exclusive scans of the per-(ring, sector) counters -> dense index
// lists + gathered feature points + per-ring offsets of the less-sharp / less-flat clouds.
struct K1bParams {
    const float4 *cloud; int N; int n_rings, NR, max_sharp, max_lsharp, max_flat;
    const int *slot_sharp, *slot_lsharp, *slot_flat; const uint8_t *slot_cnt; const int *lflat_cnt;
    int *counts; int *sharp_idx, *lsharp_idx, *flat_idx; float4 *sharp_pts, *lsharp_pts, *flat_pts;
    int cap_sharp, cap_lsharp, cap_flat; int scan_first;
    int *lsharp_ring_start, *lflat_ring_start;
    const int *ring_start; const float4 *lflat_slotted; float4 *lflat_pts;
};

// exclusive scan of one int per thread over a 256-thread block; returns the block total in `total`
__device__ __forceinline__ int k1b_block_scan(int v, int *warp_buf, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
    __syncthreads();                       // warp_buf may still be read from a previous call
    if (lane == 31) warp_buf[warp] = inc;
    __syncthreads();
    int woff = 0; total = 0;
    #pragma unroll
    for (int w = 0; w < 8; w++) { int s = warp_buf[w]; if (w < warp) woff += s; total += s; }
    return woff + inc - v;
}

__global__ void __launch_bounds__(256) k1b_compact(K1bParams p)
{
    __shared__ int off_sharp[VLO_MAX_RINGS * VLO_MAX_REGIONS + 1];
    __shared__ int off_ls[VLO_MAX_RINGS * VLO_MAX_REGIONS + 1];
    __shared__ int off_flat[VLO_MAX_RINGS * VLO_MAX_REGIONS + 1];
    __shared__ int warp_buf[8];
    __shared__ int lf_start[VLO_MAX_RINGS + 1];
    // K1B_PARTS CTAs per scan: each redoes the (tiny) prefix scans and takes its share of the copies and gathers;
    // part 0 alone publishes the counts and ring offsets
    const int b = p.scan_first + blockIdx.x, tid = threadIdx.x, part = blockIdx.y, parts = gridDim.y;
    const int nsec = p.n_rings * p.NR;
    {
        // per-thread chunk of consecutive sectors, block scan of the chunk sums, then local running offsets
        const int per = (nsec + 255) / 256, k0 = min(nsec, tid * per), k1 = min(nsec, k0 + per);
        int a = 0, l = 0, f = 0;
        for (int k = k0; k < k1; k++) { const uint8_t *c = p.slot_cnt + ((size_t)b * nsec + k) * 4; a += c[0]; l += c[1]; f += c[2]; }
        int ta, tl, tf;
        int ba = k1b_block_scan(a, warp_buf, ta), bl = k1b_block_scan(l, warp_buf, tl), bf = k1b_block_scan(f, warp_buf, tf);
        for (int k = k0; k < k1; k++) {
            const uint8_t *c = p.slot_cnt + ((size_t)b * nsec + k) * 4;
            off_sharp[k] = ba; off_ls[k] = bl; off_flat[k] = bf;
            ba += c[0]; bl += c[1]; bf += c[2];
        }
        if (tid == 0) { off_sharp[nsec] = ta; off_ls[nsec] = tl; off_flat[nsec] = tf; }
        int cnt = (tid < p.n_rings) ? p.lflat_cnt[b * p.n_rings + tid] : 0, tlf;
        int lf0 = k1b_block_scan(cnt, warp_buf, tlf);
        __syncthreads();
        if (tid <= VLO_MAX_RINGS) {
            lf_start[tid] = (tid < p.n_rings) ? lf0 : tlf;
            if (part == 0) {
                p.lsharp_ring_start[b * (VLO_MAX_RINGS + 1) + tid] = (tid < p.n_rings) ? off_ls[tid * p.NR] : tl;
                p.lflat_ring_start[b * (VLO_MAX_RINGS + 1) + tid] = (tid < p.n_rings) ? lf0 : tlf;
            }
        }
        if (tid == 0 && part == 0) { p.counts[b * 8 + 1] = ta; p.counts[b * 8 + 2] = tl; p.counts[b * 8 + 3] = tf; p.counts[b * 8 + 4] = tlf; }
    }
    __syncthreads();
    // dense less-flat cloud: ring r's centroids move from their ring slot to [lflat_ring_start[r], +cnt)
    // (coalesced float4 copies; every consumer downstream reads the dense array)
    for (int r = part * (blockDim.x >> 5) + (tid >> 5); r < p.n_rings; r += parts * (blockDim.x >> 5)) {
        int d0 = lf_start[r], s0 = p.ring_start[b * (VLO_MAX_RINGS + 1) + r];
        int cnt = p.lflat_cnt[b * p.n_rings + r];
        const float4 *src = p.lflat_slotted + (size_t)b * p.N + s0;
        float4 *dst = p.lflat_pts + (size_t)b * p.N + d0;
        for (int j = tid & 31; j < cnt; j += 32) dst[j] = src[j];
    }
    const float4 *cloud = p.cloud + (size_t)b * p.N;
    const int maxq = max(p.max_lsharp, max(p.max_sharp, p.max_flat));
    for (int k = part * blockDim.x + tid; k < nsec * maxq; k += parts * blockDim.x) {
        int sec = k / maxq, q = k % maxq;
        const uint8_t *c = p.slot_cnt + ((size_t)b * nsec + sec) * 4;
        if (q < p.max_lsharp && q < c[1]) {
            int idx = p.slot_lsharp[((size_t)b * nsec + sec) * p.max_lsharp + q];
            int o = off_ls[sec] + q;
            p.lsharp_idx[(size_t)b * p.cap_lsharp + o] = idx;
            p.lsharp_pts[(size_t)b * p.cap_lsharp + o] = cloud[idx];
        }
        if (q < p.max_sharp && q < c[0]) {
            int idx = p.slot_sharp[((size_t)b * nsec + sec) * p.max_sharp + q];
            int o = off_sharp[sec] + q;
            p.sharp_idx[(size_t)b * p.cap_sharp + o] = idx;
            p.sharp_pts[(size_t)b * p.cap_sharp + o] = cloud[idx];
        }
        if (q < p.max_flat && q < c[2]) {
            int idx = p.slot_flat[((size_t)b * nsec + sec) * p.max_flat + q];
            int o = off_flat[sec] + q;
            p.flat_idx[(size_t)b * p.cap_flat + o] = idx;
            p.flat_pts[(size_t)b * p.cap_flat + o] = cloud[idx];
        }
    }
}

int vlo_launch_extract(vlo_handle *h)
{
    ScanBatchDev &sb = h->sb;
    const vlo_config &c = h->cfg;
    K1Params p;
    p.cloud = sb.cloud; p.ring_start = sb.ring_start; p.N = c.max_points; p.n_rings = c.n_rings;
    p.K = c.curvature_region; p.NR = c.feature_regions; p.max_sharp = c.max_corner_sharp;
    p.max_lsharp = c.max_corner_less_sharp; p.max_flat = c.max_surface_flat;
    p.thr = c.surface_curvature_threshold; p.leaf = c.less_flat_filter_size;
    p.MR = c.max_ring_points; int HT = 1; while (HT < p.MR) HT <<= 1; p.HT = HT;
    p.label = sb.label; p.curvature = sb.curvature; p.picked = sb.picked;
    p.slot_sharp = sb.slot_sharp; p.slot_lsharp = sb.slot_lsharp; p.slot_flat = sb.slot_flat; p.slot_cnt = sb.slot_cnt;
    p.lflat_slotted = sb.lflat_slotted; p.lflat_cnt = sb.lflat_cnt; p.status_word = h->status_word;
    const size_t smem = k1_smem_bytes(p.MR), smem_c = k1c_smem_bytes(p.HT);
    size_t &configured = h->k1_smem_configured, &configured_c = h->k1c_smem_configured;      // per handle = per device
    if (smem > configured) {
        VLO_CUDA(cudaFuncSetAttribute(k1_extract, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    if (smem_c > configured_c) {
        VLO_CUDA(cudaFuncSetAttribute(k1c_lessflat<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        VLO_CUDA(cudaFuncSetAttribute(k1c_lessflat<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        configured_c = smem_c;
    }
    // one pass over the whole batch (sub-batches sized for the L2, so that the less-flat filter re-reads the cloud and the labels
    // from L2, were measured and lost: 0.98 ms against 0.56 ms per 128 scans; VLO_K1_SUB keeps the knob)
    const int sub = h->k1_sub > 0 ? h->k1_sub : sb.scan_count;
    vlo_prof_begin(h, ST_EXTRACT);
    for (int b0 = 0; b0 < sb.scan_count; b0 += sub) {
        const int nb = std::min(sub, sb.scan_count - b0);
        p.scan_first = sb.scan_first + b0;
        dim3 grid(c.n_rings, nb);
        k1_extract<<<grid, K1_THREADS, smem, h->stream>>>(p);
        if (p.MR <= 8 * K1_THREADS) k1c_lessflat<8><<<grid, K1_THREADS, smem_c, h->stream>>>(p);
        else k1c_lessflat<16><<<grid, K1_THREADS, smem_c, h->stream>>>(p);
        h->launches += 2;
    }
    vlo_prof_end(h, ST_EXTRACT);
    K1bParams q;
    q.cloud = sb.cloud; q.N = c.max_points; q.n_rings = c.n_rings; q.NR = c.feature_regions;
    q.max_sharp = c.max_corner_sharp; q.max_lsharp = c.max_corner_less_sharp; q.max_flat = c.max_surface_flat;
    q.slot_sharp = sb.slot_sharp; q.slot_lsharp = sb.slot_lsharp; q.slot_flat = sb.slot_flat; q.slot_cnt = sb.slot_cnt;
    q.lflat_cnt = sb.lflat_cnt; q.counts = sb.counts;
    q.sharp_idx = sb.sharp_idx; q.lsharp_idx = sb.lsharp_idx; q.flat_idx = sb.flat_idx;
    q.sharp_pts = sb.sharp_pts; q.lsharp_pts = sb.lsharp_pts; q.flat_pts = sb.flat_pts;
    q.cap_sharp = h->cap_sharp; q.cap_lsharp = h->cap_lsharp; q.cap_flat = h->cap_flat;
    q.lsharp_ring_start = sb.lsharp_ring_start; q.lflat_ring_start = sb.lflat_ring_start;
    q.ring_start = sb.ring_start; q.lflat_slotted = sb.lflat_slotted; q.lflat_pts = sb.lflat_pts;
    q.scan_first = sb.scan_first;
    VLO_PROF(h, ST_COMPACT, (k1b_compact<<<dim3(sb.scan_count, sb.scan_count >= 64 ? 8 : 16), 256, 0, h->stream>>>(q)));
    h->launches += 1;
    VLO_CUDA(cudaGetLastError());
    return VLO_OK;
}
