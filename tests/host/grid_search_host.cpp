// TEST INFRASTRUCTURE -- host emulation of the thread-level voxel-hash searches of vil_sensor_fusion_b200/csrc/grid.cuh.
// The header is compiled as-is for the CPU (g++, -ffp-contract=off like nvcc --fmad=false): the CUDA qualifiers come
// from <cuda_runtime.h>, the handful of device intrinsics the thread-level code uses are defined below, and the warp
// collectives (used only by functions this harness never calls) are declared so that the header parses.
// What it checks (tests/test_host_grid_search.py): grid_search_thread27 and grid_search_thread return exactly the
// (d2, tie)-lexicographic top-K of a brute-force scan -- K = 5 plain (scan-to-map) and K = 1 with the partner filter
// (scan-to-scan) -- on clouds with duplicates and lattice ties, for queries inside, at the border of and outside the
// populated cells, with and without a caller-supplied bound.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
using std::isfinite;                      // CUDA's global overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
// warp collectives: never executed here
template <class T> T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> T __shfl_up_sync(unsigned, T v, int) { return v; }
template <class T> T __shfl_down_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return v; }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return v; }
static inline int __reduce_min_sync(unsigned, int v) { return v; }
static inline int __reduce_max_sync(unsigned, int v) { return v; }
static inline int __reduce_add_sync(unsigned, int v) { return v; }
static inline unsigned __match_any_sync(unsigned, int) { return 1u; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __any_sync(unsigned, int p) { return p; }
static inline int __all_sync(unsigned, int p) { return p; }

#include "../../vil_sensor_fusion_b200/csrc/grid.cuh"
#include "../../vil_sensor_fusion_b200/csrc/dense6.cuh"
#include "../../vil_sensor_fusion_b200/csrc/map_lin.cuh"
#include "../../vil_sensor_fusion_b200/csrc/odom_lin.cuh"
extern "C" {
#include "../../oracle/detmath.h"
}

// ---- host build of one grid (same layout contract as k2_grid.cu: keys / start / sorted, tag = ring << 24 | dense index)
struct HostGrid {
    GridSet gs;
    std::vector<unsigned long long> keys; std::vector<int> cnt, start; std::vector<float4> sorted;
};
static void build_grid(HostGrid &g, const float *pts, const int *ring, int n, float cell)
{
    int ts = 1024; while (ts < 4 * n) ts <<= 1;
    g.keys.assign(ts, GRID_EMPTY); g.cnt.assign(ts, 0); g.start.assign(ts + 1, 0); g.sorted.resize(n > 0 ? n : 1);
    g.gs.cell = cell; g.gs.inv_cell = 1.0f / cell; g.gs.ts = ts; g.gs.max_pts = n; g.gs.G = 1;
    std::vector<int> slot_of(n);
    for (int i = 0; i < n; i++) {
        int ix = (int)floorf(pts[4 * i] * g.gs.inv_cell), iy = (int)floorf(pts[4 * i + 1] * g.gs.inv_cell), iz = (int)floorf(pts[4 * i + 2] * g.gs.inv_cell);
        unsigned long long key = grid_key(ix, iy, iz);
        int slot = (int)(grid_hash(ix, iy, iz) & (unsigned)(ts - 1));
        while (g.keys[slot] != GRID_EMPTY && g.keys[slot] != key) slot = (slot + 1) & (ts - 1);
        g.keys[slot] = key; g.cnt[slot]++; slot_of[i] = slot;
    }
    for (int s = 0; s < ts; s++) g.start[s + 1] = g.start[s] + g.cnt[s];
    std::vector<int> cur(g.start.begin(), g.start.end() - 1);
    // storage order inside a cell must not matter: fill back to front
    for (int i = n - 1; i >= 0; i--) {
        unsigned tag = ((unsigned)ring[i] << 24) | (unsigned)i;
        g.sorted[cur[slot_of[i]]++] = make_float4(pts[4 * i], pts[4 * i + 1], pts[4 * i + 2], __uint_as_float(tag));
    }
    g.gs.keys = g.keys.data(); g.gs.cnt = g.cnt.data(); g.gs.start = g.start.data(); g.gs.sorted = g.sorted.data(); g.gs.bsum = nullptr;
}

extern "C" {
// plain K-NN (FilterAll), K = 5: mode 0 = grid_search_thread27 (cell must be >= sqrt(dmax)), 1 = grid_search_thread (any cell),
// bound_mode 1: pass the exact K-th distance of a brute-force pass as the caller's bound (what k5 does from iteration 1 on)
int host_knn5(const float *pts, int n, float cell, const float *q, int nq, float dmax, int mode, int bound_mode, int *idx_out, float *d2_out)
{
    std::vector<int> ring(n, 0);
    HostGrid g; build_grid(g, pts, ring.data(), n, cell);
    for (int k = 0; k < nq; k++) {
        float qx = q[3 * k], qy = q[3 * k + 1], qz = q[3 * k + 2];
        float bound = -1.0f;
        if (bound_mode) {
            std::vector<float> d;
            for (int i = 0; i < n; i++) { float dx = pts[4 * i] - qx, dy = pts[4 * i + 1] - qy, dz = pts[4 * i + 2] - qz; float d2 = (dx * dx + dy * dy) + dz * dz; if (d2 < dmax) d.push_back(d2); }
            if (d.size() >= 5) { std::sort(d.begin(), d.end()); bound = d[4]; }
        }
        TopKI<5> best;
        if (mode == 0) grid_search_thread27(g.gs, 0, qx, qy, qz, dmax, FilterAll(), best, bound);
        else grid_search_thread(g.gs, 0, qx, qy, qz, dmax, grid_thread_rho(cell, dmax), FilterAll(), best, bound);
        for (int j = 0; j < 5; j++) {
            idx_out[5 * k + j] = best.valid(j) ? best.index(j) : -1;
            d2_out[5 * k + j] = __uint_as_float((unsigned)(best.key[j] >> 32));
        }
    }
    return 0;
}
// filtered 1-NN with upstream's partner rules (FilterPartner): the 27-cell search limited to one cell edge, as
// k3_assoc_thread uses it; returns -2 where that search cannot decide (the kernel falls back to the exact warp search)
int host_partner(const float *pts, const int *ring, int n, float cell, const float *q, int nq, const int *ind, const int *ring_lo,
                 const int *ring_hi, const int *skip, int fwd_bound, int *idx_out)
{
    HostGrid g; build_grid(g, pts, ring, n, cell);
    const float edge = cell - 2e-3f * cell, dfast = fminf(25.0f, edge * edge);
    for (int k = 0; k < nq; k++) {
        FilterPartner f; f.ind = ind[k]; f.ring_lo = ring_lo[k]; f.ring_hi = ring_hi[k]; f.skip_ring = skip[k]; f.fwd_bound = fwd_bound;
        TopKT<1> best;
        grid_search_thread27(g.gs, 0, q[3 * k], q[3 * k + 1], q[3 * k + 2], dfast, f, best);
        idx_out[k] = best.tag[0] != GRID_NOTAG ? (int)(best.tag[0] & 0xFFFFFFu) : (dfast < 25.0f ? -2 : -1);
    }
    return 0;
}

// D1 of DESIGN.md: the product's deterministic sin/cos/atan/atan2 (vlo_internal.cuh) and the oracle's (oracle/detmath.h) are
// two texts of the same operation sequence -- count the inputs on which they differ in any bit
static inline bool same_bits(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }
long host_detmath_mismatches(const float *x, const float *y, long n)
{
    long bad = 0;
    for (long i = 0; i < n; i++) {
        float s1, c1, s2, c2;
        vlo_sincosf(x[i], s1, c1); orc_sincosf(x[i], &s2, &c2);
        if (!same_bits(s1, s2) || !same_bits(c1, c2)) bad++;
        if (!same_bits(vlo_atanf(x[i]), orc_atanf(x[i]))) bad++;
        if (!same_bits(vlo_atan2f(y[i], x[i]), orc_atan2f(y[i], x[i]))) bad++;
    }
    return bad;
}

// the scan-to-map linearisation of csrc/map_lin.cuh and the 6x6 solve of csrc/dense6.cuh, one call per input set
void host_eig3(const float *A9, float *eval3, float *evec9) { eig3_jacobi(A9, eval3, evec9); }
void host_lstsq53(const float *A15, float *x3) { lstsq53(reinterpret_cast<const float (*)[3]>(A15), x3); }
int host_edge_coeff(const float *sel4, const float *nb20, float *coeff4)
{
    float4 nb[5];
    for (int j = 0; j < 5; j++) nb[j] = make_float4(nb20[4 * j], nb20[4 * j + 1], nb20[4 * j + 2], nb20[4 * j + 3]);
    return map_edge_coeff(make_float4(sel4[0], sel4[1], sel4[2], sel4[3]), nb, coeff4) ? 1 : 0;
}
int host_plane_coeff(const float *sel4, const float *nb20, float *coeff4)
{
    float4 nb[5];
    for (int j = 0; j < 5; j++) nb[j] = make_float4(nb20[4 * j], nb20[4 * j + 1], nb20[4 * j + 2], nb20[4 * j + 3]);
    return map_plane_coeff(make_float4(sel4[0], sel4[1], sel4[2], sel4[3]), nb, coeff4) ? 1 : 0;
}
void host_solve6(const float *A36, const float *b6, float *x6) { vlo_solve6_colpiv_qr(A36, b6, x6); }
void host_to_map(const float *T6, const float *p4, float *out4)
{
    float trig[6];
    for (int a = 0; a < 3; a++) vlo_sincosf(T6[a], trig[2 * a], trig[2 * a + 1]);
    float4 o = to_map(T6, trig, make_float4(p4[0], p4[1], p4[2], p4[3]));
    out4[0] = o.x; out4[1] = o.y; out4[2] = o.z; out4[3] = o.w;
}

// the scan-to-scan linearisation of csrc/odom_lin.cuh and transformToStart of vlo_internal.cuh
static inline float4 f4(const float *p) { return make_float4(p[0], p[1], p[2], p[3]); }
int host_odom_edge_coeff(const float *sel, const float *a, const float *b, int iter, float *coeff4) { return edge_coeff(f4(sel), f4(a), f4(b), iter, coeff4) ? 1 : 0; }
int host_odom_plane_coeff(const float *sel, const float *t1, const float *t2, const float *t3, int iter, float *coeff4)
{ return plane_coeff(f4(sel), f4(t1), f4(t2), f4(t3), iter, coeff4) ? 1 : 0; }
void host_odom_jacobian_row(const float *T6, const float *ori4, const float *coeff4, float *row6, float *bval)
{
    float trig[6];
    for (int a = 0; a < 3; a++) vlo_sincosf(T6[a], trig[2 * a], trig[2 * a + 1]);
    odom_jacobian_row(T6, trig, ori4[0], ori4[1], ori4[2], coeff4, row6, *bval);
}
void host_to_start(const float *T6, const float *p4, int deskew, float inv_period, float *out4)
{
    float4 o = vlo_to_start(T6, f4(p4), deskew, inv_period);
    out4[0] = o.x; out4[1] = o.y; out4[2] = o.z; out4[3] = o.w;
}
}
