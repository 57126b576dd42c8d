// Voxel-hash grids (K2): device-side layout and the exact, warp-cooperative nearest-neighbour
// search shared by scan-to-scan association (k3) and scan-to-map (k5).  Replaces
// pcl::KdTreeFLANN::nearestKSearch as used by the `loam` nodelets.
//
// Layout (all grids of a set are contiguous so they clear with two memsets):
//   keys  [G][ts]   packed cell coordinates (3 x 21 bit, biased), ~0 = empty, open addressing
//   cnt   [G][ts]   points per cell during the build (counted down to 0 by the scatter pass)
//   start [G][ts+1] exclusive scan of cnt in slot order: cell s owns sorted[start[s] .. start[s+1])
//   sorted[G][max]  xyz + tag; tag = (ring << 24) | dense index of the point in its ring-major cloud
// Exactness contract: neighbours compare by (d2, tie) lexicographically with
// d2 = ((dx*dx) + (dy*dy)) + (dz*dz) in float32 (no FMA); shells of cells are visited outwards and
// the search stops only when the k-th best distance is strictly inside the visited cube (minus a
// slack that covers the rounding of floor(p * inv_cell)), so storage order never matters.
//
// Work distribution (round-1 ncu: the lane-per-cell version spent 30 % of its instructions in
// divergent top-k insertions): the 32 lanes first resolve up to 32 cells of the shell (one hash
// probe each), a warp scan turns the cell populations into one flat candidate range, and the lanes
// then stride over that range -- every lane sees the same number of candidates whatever the
// per-cell populations are.
#pragma once
#include "vlo_internal.cuh"

#define GRID_EMPTY 0xFFFFFFFFFFFFFFFFull
#define GRID_SCRATCH_INTS 64      // per-warp shared-memory scratch (inclusive prefix + base offsets)

__device__ __forceinline__ unsigned long long grid_key(int ix, int iy, int iz)
{
    return ((unsigned long long)(unsigned)(ix + (1 << 20)) << 42) | ((unsigned long long)(unsigned)(iy + (1 << 20)) << 21)
         | (unsigned long long)(unsigned)(iz + (1 << 20));
}
__device__ __forceinline__ unsigned grid_hash(int ix, int iy, int iz)
{
    return ((unsigned)ix * 73856093u) ^ ((unsigned)iy * 19349663u) ^ ((unsigned)iz * 83492791u);
}

// returns slot of the cell or -1
__device__ __forceinline__ int grid_find(const GridSet &gs, int g, int ix, int iy, int iz)
{
    const unsigned long long *keys = gs.keys + (size_t)g * gs.ts;
    unsigned long long key = grid_key(ix, iy, iz);
    int slot = (int)(grid_hash(ix, iy, iz) & (unsigned)(gs.ts - 1));
    while (true) {
        unsigned long long k = keys[slot];
        if (k == key) return slot;
        if (k == GRID_EMPTY) return -1;
        slot = (slot + 1) & (gs.ts - 1);
    }
}

// Candidate filters -------------------------------------------------------------------------------
struct FilterAll {          // plain k-NN, ties -> lowest dense index
    __device__ __forceinline__ bool operator()(unsigned tag, unsigned &tie) const { tie = tag & 0xFFFFFFu; return true; }
};
// upstream's partner loops: forward indices first (ascending), then backward
// (descending); `ring_lo..ring_hi` admissible rings, `skip_ring` excluded (or -1).
struct FilterPartner {
    int ind;        // dense index of the nearest neighbour
    int ring_lo, ring_hi, skip_ring;
    int fwd_bound;  // forward candidates need dense index < fwd_bound
    __device__ __forceinline__ bool operator()(unsigned tag, unsigned &tie) const
    {
        int ring = (int)(tag >> 24), idx = (int)(tag & 0xFFFFFFu);
        if (ring < ring_lo || ring > ring_hi || ring == skip_ring || idx == ind) return false;
        if (idx > ind) { if (idx >= fwd_bound) return false; tie = (unsigned)(idx - ind); }
        else tie = 0x40000000u + (unsigned)(ind - idx);
        return true;
    }
};

// upstream's partner loops in full (oracle/laser_odometry.c partners_corner / partners_surf; the same rule as SegFilter mode 1,
// segbox.cuh): around the nearest neighbour `ind` whose scan id is `scan` the forward loop visits the dense indices (ind, hi),
// the backward loop (lo, ind) -- lo / hi from k3_partner_ranges -- and a forward point is in the same-scan class (want 2) if
// its scan id (tag >> 24 = int(w)) is <= scan, a backward point if it is >= scan, otherwise in the other-scan class (want 3)
struct FilterPartnerRange {
    int ind, scan, lo, hi, want;
    __device__ __forceinline__ bool operator()(unsigned tag, unsigned &tie) const
    {
        const int e = (int)(tag >> 24), idx = (int)(tag & 0xFFFFFFu);
        if (idx <= lo || idx >= hi || idx == ind) return false;
        int cls;
        if (idx > ind) { cls = e > scan ? 3 : 2; tie = (unsigned)(idx - ind); }
        else { cls = e < scan ? 3 : 2; tie = 0x40000000u + (unsigned)(ind - idx); }
        return cls == want;
    }
};

// K-best list kept sorted ascending by (d2 bits, tie) in registers.  Entries whose tag is
// GRID_NOTAG are placeholders carrying the admission threshold; they always sit behind real ones.
#define GRID_NOTAG 0xFFFFFFFFu
template <int K> struct TopK {
    unsigned d[K]; unsigned t[K]; unsigned tag[K];
    __device__ __forceinline__ void init(unsigned dthr, unsigned tthr) {
        #pragma unroll
        for (int i = 0; i < K; i++) { d[i] = dthr; t[i] = tthr; tag[i] = GRID_NOTAG; }
    }
    __device__ __forceinline__ void insert(unsigned dd, unsigned tt, unsigned tg) {
        if (!(dd < d[K - 1] || (dd == d[K - 1] && tt < t[K - 1]))) return;
        // walk down from the end shifting larger entries up; stop at the first smaller one
        bool placed = false;
        #pragma unroll
        for (int i = K - 1; i > 0; i--) {
            if (!placed) {
                bool up = dd < d[i - 1] || (dd == d[i - 1] && tt < t[i - 1]);
                if (up) { d[i] = d[i - 1]; t[i] = t[i - 1]; tag[i] = tag[i - 1]; }
                else { d[i] = dd; t[i] = tt; tag[i] = tg; placed = true; }
            }
        }
        if (!placed) { d[0] = dd; t[0] = tt; tag[0] = tg; }
    }
    __device__ __forceinline__ void pop() {
        #pragma unroll
        for (int i = 0; i < K - 1; i++) { d[i] = d[i + 1]; t[i] = t[i + 1]; tag[i] = tag[i + 1]; }
        d[K - 1] = 0xFFFFFFFFu; t[K - 1] = 0xFFFFFFFFu; tag[K - 1] = GRID_NOTAG;
    }
};

// Warp-cooperative exact search.  All 32 lanes call it with the same query; on return every lane
// holds the global K best in `best` (d = float bits of d2; tag == GRID_NOTAG means "none").
// Candidates need d2 < dmax (strict) and must pass the filter.  `scratch`: GRID_SCRATCH_INTS ints of
// shared memory private to this warp.
template <int K, typename Filter>
__device__ void grid_search(const GridSet &gs, int g, float qx, float qy, float qz, float dmax,
                            const Filter &flt, TopK<K> &best, int lane, int *scratch)
{
    const int *start = gs.start + (size_t)g * (gs.ts + 1);
    const float4 *sorted = gs.sorted + (size_t)g * gs.max_pts;
    const int cx = (int)floorf(qx * gs.inv_cell), cy = (int)floorf(qy * gs.inv_cell), cz = (int)floorf(qz * gs.inv_cell);
    best.init(__float_as_uint(dmax), 0xFFFFFFFFu);
    const float slack = 1e-3f * gs.cell;
    for (int rho = 1; ; rho++) {
        const int side = 2 * rho + 1, ncell = side * side * side;   // rho == 1 also covers the centre cell
        TopK<K> loc;
        if (lane == 0) loc = best; else loc.init(best.d[K - 1], best.t[K - 1]);
        for (int c0 = 0; c0 < ncell; c0 += 32) {
            const int c = c0 + lane;
            int s0 = 0, n = 0;
            if (c < ncell) {
                int dx, dy, dz;
                if (rho == 1) { dx = c % 3 - 1; dy = (c / 3) % 3 - 1; dz = c / 9 - 1; }            // constant divisors
                else { dx = c % side - rho; dy = (c / side) % side - rho; dz = c / (side * side) - rho; }
                bool shell = rho == 1 || abs(dx) == rho || abs(dy) == rho || abs(dz) == rho;   // interior: already visited
                if (shell) {
                    int slot = grid_find(gs, g, cx + dx, cy + dy, cz + dz);
                    if (slot >= 0) { s0 = start[slot]; n = start[slot + 1] - s0; }
                }
            }
            int inc = n;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
            const int total = __shfl_sync(0xffffffffu, inc, 31);
            if (total == 0) continue;
            scratch[lane] = inc;                 // inclusive prefix of the cell populations
            scratch[32 + lane] = s0 - (inc - n); // flat index j inside cell e maps to sorted[base[e] + j]
            __syncwarp();
            for (int j = lane; j < total; j += 32) {
                int e = 0;                       // first cell whose inclusive prefix exceeds j
                #pragma unroll
                for (int step = 16; step >= 1; step >>= 1) if (scratch[e + step - 1] <= j) e += step;
                float4 p = sorted[scratch[32 + e] + j];
                float ddx = p.x - qx, ddy = p.y - qy, ddz = p.z - qz;
                float d2 = (ddx * ddx + ddy * ddy) + ddz * ddz;
                unsigned tie, tag = __float_as_uint(p.w);
                if (!(d2 < dmax) || !flt(tag, tie)) continue;
                loc.insert(__float_as_uint(d2), tie, tag);
            }
            __syncwarp();
        }
        #pragma unroll
        for (int i = 0; i < K; i++) {
            bool valid = loc.tag[0] != GRID_NOTAG;
            unsigned hd = valid ? loc.d[0] : 0xFFFFFFFFu;
            unsigned m = __reduce_min_sync(0xffffffffu, hd);
            if (m == 0xFFFFFFFFu) { best.d[i] = __float_as_uint(dmax); best.t[i] = 0xFFFFFFFFu; best.tag[i] = GRID_NOTAG; continue; }
            unsigned ht = (valid && hd == m) ? loc.t[0] : 0xFFFFFFFFu;
            unsigned mt = __reduce_min_sync(0xffffffffu, ht);
            bool mine = valid && hd == m && loc.t[0] == mt;
            int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
            unsigned tg = __shfl_sync(0xffffffffu, loc.tag[0], src);
            best.d[i] = m; best.t[i] = mt; best.tag[i] = tg;
            if (lane == src) loc.pop();
        }
        // stop when the k-th best is strictly inside the visited cube, or nothing below dmax can remain
        float bound = (float)rho * gs.cell - slack;
        float b2 = bound * bound;
        if (best.tag[K - 1] != GRID_NOTAG && __uint_as_float(best.d[K - 1]) < b2) break;
        if (b2 >= dmax) break;
    }
}

// ----------------------------------------------------------------------------------------------
// Thread-level exact search (one query per thread) for searches whose radius is about one cell:
// scan-to-map 5-NN with d2 < 1 on ~1 m cells.  Round-1 ncu of the warp-cooperative version on this
// workload: 540 warp-instructions per query, issue-bound, most of it per-query overhead (27 hash
// probes, prefix scan, per-candidate binary search, 5-round merge).  Here a thread walks the
// (2 rho + 1)^3 cells around its query nearest-first and skips every cell (and whole rows / slabs)
// whose box cannot hold a point closer than the current K-th best, so a typical query probes ~8 cells
// and reads ~50 candidates.  Neighbouring queries (consecutive points of a ring) visit the same cells,
// which keeps the loads in L1/L2.  Same exactness contract as grid_search: (d2, tie) lexicographic,
// d2 = ((dx*dx) + (dy*dy)) + (dz*dz), pruning bounds shrunk by `slack` to cover floor() rounding.
template <int K> struct TopKT {
    unsigned long long key[K];      // (float bits of d2) << 32 | tie
    unsigned tag[K];
    __device__ __forceinline__ void init(float dmax) {
        #pragma unroll
        for (int i = 0; i < K; i++) { key[i] = (unsigned long long)__float_as_uint(dmax) << 32; tag[i] = GRID_NOTAG; }
    }
    __device__ __forceinline__ float kth() const { return __uint_as_float((unsigned)(key[K - 1] >> 32)); }
    __device__ __forceinline__ void insert(unsigned long long kk, unsigned tg) {
        if (kk >= key[K - 1]) return;
        #pragma unroll
        for (int i = K - 1; i > 0; i--) {
            bool a = kk < key[i - 1], b = kk < key[i];
            key[i] = a ? key[i - 1] : (b ? kk : key[i]);
            tag[i] = a ? tag[i - 1] : (b ? tg : tag[i]);
        }
        if (kk < key[0]) { key[0] = kk; tag[0] = tg; }
    }
};

// key-only variant for plain k-NN (FilterAll: the tie IS the dense index, so no tag is carried)
template <int K> struct TopKI {
    unsigned long long key[K];      // (float bits of d2) << 32 | dense index; low word ~0 = empty
    // placeholders admit every key with d2 <= thr (the search pre-filters d2 < dmax, so thr = dmax stays exclusive)
    __device__ __forceinline__ void init(float thr) {
        #pragma unroll
        for (int i = 0; i < K; i++) key[i] = ((unsigned long long)__float_as_uint(thr) << 32) | 0xFFFFFFFFull;
    }
    __device__ __forceinline__ float kth() const { return __uint_as_float((unsigned)(key[K - 1] >> 32)); }
    __device__ __forceinline__ bool valid(int i) const { return (unsigned)key[i] != 0xFFFFFFFFu; }
    __device__ __forceinline__ int index(int i) const { return (int)((unsigned)key[i] & 0xFFFFFFu); }
    __device__ __forceinline__ void insert(unsigned long long kk, unsigned) {
        if (kk >= key[K - 1]) return;
        #pragma unroll
        for (int i = 0; i < K; i++) {           // bubble the new key down, carrying the displaced one
            bool lt = kk < key[i];
            unsigned long long lo = lt ? kk : key[i];
            kk = lt ? key[i] : kk;
            key[i] = lo;
        }
    }
};

// offset of step a (0, 1, 2, 3, 4, ..) along an axis: 0, +s, -s, +2s, -2s, ..  (s = side of the nearer face)
__device__ __forceinline__ int grid_step_offset(int a, int s) { int m = (a + 1) >> 1; return (a & 1) ? s * m : -s * m; }
// lower bound of |q - p| along one axis for points binned `o` cells away from the query's cell c
__device__ __forceinline__ float grid_axis_lb(int o, int c, float q, float cell, float slack)
{
    if (o == 0) return 0.0f;
    float d = (o > 0) ? ((float)(c + o) * cell - q) - slack : (q - (float)(c + o + 1) * cell) - slack;
    return fmaxf(d, 0.0f);
}

// `bound` < dmax (optional): a distance known to be reached by at least K admissible points (e.g. the previous
// iteration's neighbours seen from the new pose) -- the walk then starts with a tight pruning radius.
template <typename Top, typename Filter>
__device__ __forceinline__ void grid_search_thread(const GridSet &gs, int g, float qx, float qy, float qz, float dmax, int rho,
                                                   const Filter &flt, Top &best, float bound = -1.0f)
{
    const int *start = gs.start + (size_t)g * (gs.ts + 1);
    const float4 *sorted = gs.sorted + (size_t)g * gs.max_pts;
    const float cell = gs.cell, slack = 1e-3f * gs.cell;
    const float fx = qx * gs.inv_cell, fy = qy * gs.inv_cell, fz = qz * gs.inv_cell;
    const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
    const int sx = (fx - (float)cx >= 0.5f) ? 1 : -1, sy = (fy - (float)cy >= 0.5f) ? 1 : -1, sz = (fz - (float)cz >= 0.5f) ? 1 : -1;
    const int side = 2 * rho + 1;
    best.init((bound >= 0.0f && bound < dmax) ? bound : dmax);
    for (int az = 0; az < side; az++) {
        const int oz = grid_step_offset(az, sz);
        const float lz = grid_axis_lb(oz, cz, qz, cell, slack), lz2 = lz * lz;
        if (lz2 > best.kth()) continue;
        for (int ay = 0; ay < side; ay++) {
            const int oy = grid_step_offset(ay, sy);
            const float ly = grid_axis_lb(oy, cy, qy, cell, slack), lyz2 = ly * ly + lz2;
            if (lyz2 > best.kth()) continue;
            for (int ax = 0; ax < side; ax++) {
                const int ox = grid_step_offset(ax, sx);
                const float lx = grid_axis_lb(ox, cx, qx, cell, slack);
                if (lx * lx + lyz2 > best.kth()) continue;
                const int slot = grid_find(gs, g, cx + ox, cy + oy, cz + oz);
                if (slot < 0) continue;
                const int j0 = start[slot], j1 = start[slot + 1];
                for (int j = j0; j < j1; j++) {
                    const float4 p = sorted[j];
                    const float ddx = p.x - qx, ddy = p.y - qy, ddz = p.z - qz;
                    const float d2 = (ddx * ddx + ddy * ddy) + ddz * ddz;
                    unsigned tie; const unsigned tag = __float_as_uint(p.w);
                    if (!(d2 < dmax) || !flt(tag, tie)) continue;
                    best.insert(((unsigned long long)__float_as_uint(d2) << 32) | tie, tag);
                }
            }
        }
    }
}

// rho == 1 specialisation (the scan-to-map case: d2 < 1 on ~1 m cells).  The nested walk above makes a warp step
// through the UNION of the cells its lanes visit (ncu r01c: 13.6 of 32 lanes active on average); here every lane keeps
// its own 27-bit work mask and pops its own next cell, so the warp iterates max-over-lanes(non-empty cells) times.
//   1. own cell (all lanes together): gives a k-th distance that is already nearly final
//   2. one unrolled pass marks the neighbour cells whose box can still hold a closer point
//   3. pop loop: next marked cell in nearest-first order, re-test against the current k-th distance, hash probe;
//      pruned / absent / empty cells are skipped inside the pop, the candidate loop runs only on populated cells
// Bit c of the mask is cell (ex, ey, ez), e = 0 own / 1 neighbour on the nearer side / 2 on the farther side, packed
// two bits per cell in G27_E*; order = (far sides, non-zero offsets) ascending.  Same exactness contract as above: a
// cell is skipped only when its box lower bound exceeds the k-th best, so the result is the exact (d2, tie) top-K.
#define G27_EX 0x268a5849824504ull
#define G27_EY 0x29a26522485110ull
#define G27_EZ 0x2a689694205440ull
#define G27_REBUILDS 3           // mask passes: after the own cell and after each of the next two scanned cells

__device__ __forceinline__ float grid_sel3(float a1, float a2, int e) { return e == 0 ? 0.0f : (e == 1 ? a1 : a2); }

template <typename Top, typename Filter>
__device__ __forceinline__ void grid_scan_range(const float4 *sorted, int j0, int j1, float qx, float qy, float qz, float dmax,
                                                const Filter &flt, Top &best)
{
    // four candidates per round, their loads issued together (the r01c profile had 14 % of the stall samples on the
    // one-load-per-iteration dependency); indices past the end are clamped to the last point and not inserted
    for (int j = j0; j < j1; j += 4) {
        float4 p[4];
        #pragma unroll
        for (int u = 0; u < 4; u++) p[u] = sorted[min(j + u, j1 - 1)];
        #pragma unroll
        for (int u = 0; u < 4; u++) {
            // the filter looks at the tag only: candidates it rejects (four of five in a partner search) cost no arithmetic
            unsigned tie; const unsigned tag = __float_as_uint(p[u].w);
            if (j + u < j1 && flt(tag, tie)) {
                const float ddx = p[u].x - qx, ddy = p[u].y - qy, ddz = p[u].z - qz;
                const float d2 = (ddx * ddx + ddy * ddy) + ddz * ddz;
                if (d2 < dmax) best.insert(((unsigned long long)__float_as_uint(d2) << 32) | tie, tag);
            }
        }
    }
}

template <typename Top, typename Filter>
__device__ __forceinline__ void grid_search_thread27(const GridSet &gs, int g, float qx, float qy, float qz, float dmax,
                                                     const Filter &flt, Top &best, float bound = -1.0f)
{
    const int *start = gs.start + (size_t)g * (gs.ts + 1);
    const float4 *sorted = gs.sorted + (size_t)g * gs.max_pts;
    const float cell = gs.cell, slack = 1e-3f * gs.cell;
    const float fx = qx * gs.inv_cell, fy = qy * gs.inv_cell, fz = qz * gs.inv_cell;
    const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
    const int sx = (fx - (float)cx >= 0.5f) ? 1 : -1, sy = (fy - (float)cy >= 0.5f) ? 1 : -1, sz = (fz - (float)cz >= 0.5f) ? 1 : -1;
    best.init((bound >= 0.0f && bound < dmax) ? bound : dmax);
    // squared per-axis lower bounds of the nearer ([1]) and farther ([2]) neighbour cell; [0] (own cell) is 0
    float lx[3], ly[3], lz[3];
    lx[0] = 0.0f; ly[0] = 0.0f; lz[0] = 0.0f;
    { float a = grid_axis_lb(sx, cx, qx, cell, slack), b = grid_axis_lb(-sx, cx, qx, cell, slack); lx[1] = a * a; lx[2] = b * b; }
    { float a = grid_axis_lb(sy, cy, qy, cell, slack), b = grid_axis_lb(-sy, cy, qy, cell, slack); ly[1] = a * a; ly[2] = b * b; }
    { float a = grid_axis_lb(sz, cz, qz, cell, slack), b = grid_axis_lb(-sz, cz, qz, cell, slack); lz[1] = a * a; lz[2] = b * b; }
    // one loop body serves the own cell (bit 0, popped by every lane in the first round) and the neighbours: the
    // mask of neighbour cells is built right after the own cell has been scanned (single copy of the scan code --
    // the kernel's instruction footprint matters: ncu showed no_instruction stalls with 150 KB of SASS)
    unsigned mask = 1u;
    int rebuild = G27_REBUILDS;
    while (mask) {
        int j0 = 0, j1 = 0;
        do {
            const int c = __ffs(mask) - 1;
            mask &= mask - 1u;
            const int ex = (int)(unsigned)(G27_EX >> (2 * c)) & 3, ey = (int)(unsigned)(G27_EY >> (2 * c)) & 3, ez = (int)(unsigned)(G27_EZ >> (2 * c)) & 3;
            const float l2 = grid_sel3(lx[1], lx[2], ex) + (grid_sel3(ly[1], ly[2], ey) + grid_sel3(lz[1], lz[2], ez));
            if (l2 > best.kth()) continue;
            const int ox = ex == 0 ? 0 : (ex == 1 ? sx : -sx), oy = ey == 0 ? 0 : (ey == 1 ? sy : -sy), oz = ez == 0 ? 0 : (ez == 1 ? sz : -sz);
            const int slot = grid_find(gs, g, cx + ox, cy + oy, cz + oz);
            if (slot < 0) continue;
            j0 = start[slot]; j1 = start[slot + 1];
        } while (j1 == j0 && mask);
        grid_scan_range(sorted, j0, j1, qx, qy, qz, dmax, flt, best);
        if (rebuild > 0) {
            // after the own cell: mark the neighbour cells that can still hold a closer point; after each of the next
            // scans: drop the marks the tighter k-th distance has made pointless (a lane whose own cell held fewer than K
            // points starts with all 26 neighbours marked and would otherwise pop and re-test every one of them while
            // the other 31 lanes wait)
            const unsigned keep = (rebuild == G27_REBUILDS) ? 0xFFFFFFFFu : mask;
            rebuild--;
            const float kth = best.kth();
            unsigned m2 = 0u;
            #pragma unroll
            for (int c = 1; c < 27; c++) {
                const int ex = (int)((G27_EX >> (2 * c)) & 3ull), ey = (int)((G27_EY >> (2 * c)) & 3ull), ez = (int)((G27_EZ >> (2 * c)) & 3ull);
                if (!(lx[ex] + (ly[ey] + lz[ez]) > kth)) m2 |= 1u << c;
            }
            mask = m2 & keep;
        }
    }
}

// cells per axis side the thread search must cover so that no point with d2 < dmax is missed
static inline int grid_thread_rho(float cell, float dmax)
{
    float need = sqrtf(dmax) + 1e-3f * cell;
    int rho = (int)ceilf(need / cell);
    return rho < 1 ? 1 : rho;
}
