// Ring-segment box index: the spatial index of the scan-to-scan association (LaserOdometry, k3_odometry.cu).
// Replaces pcl::KdTreeFLANN::nearestKSearch + the ring-constrained partner loops of BasicLaserOdometry::process in the
// `loam` nodelet laserOdometry (gtsam_fusion/launch/loam.launch:40-45).
//
// The target clouds of a sweep (less-sharp corners, less-flat surface points) are ring-major arrays: ring r owns the
// dense indices [ring_start[r], ring_start[r + 1]), consecutive points of a ring are neighbours in space.  The index
// keeps that order (no sorted copy, the dense index IS the position) and adds two levels of axis-aligned boxes:
//   fine   segment f = up to SEG_PTS consecutive points of ONE ring (a short arc: its box is tight); ring r owns segments
//                      [seg_ring[r], seg_ring[r + 1])
//   coarse group c   = 32 fine segments that are close in space: the fine segments are counting-sorted by the cell of
//                      their box centre on a 16 x 4 x 16 grid over the cloud's bounding box (Morton order); `mbox`
//                      holds the fine boxes in that order, group c = mbox[32 c .. 32 c + 32)
// A query is answered by ONE WARP: every level is one box test per lane, a ballot, and a 32-wide coalesced scan of
// the surviving segments -- no divergence, no hash probes, no data-dependent shell expansion (round-1 ncu of the
// voxel-hash version: 8.9 of 32 lanes active, 13 % of the samples on the probe load, far partners falling back to a
// cooperative search that walked thousands of empty cells).  Ring-constrained partner searches use the ring-major
// numbering directly: the admissible rings' segments are one contiguous range.
//
// Exactness: a box is skipped only if its lower bound exceeds the best distance found so far.  The bound is computed
// with the same rounded operations as the point distance -- per-axis gap g = max(lo - q, q - hi, 0) satisfies
// fl(g) <= fl(|p - q|) for every p in the box because IEEE subtraction is monotone, and so do the squares and the
// ((x + y) + z) sums -- so lb <= d2 holds in float32, ties included (boxes with lb == best are visited).  Candidates
// compare by (d2 bits, tie) lexicographically like every arg-min of this library.
#pragma once
#include "vlo_internal.cuh"

#ifndef SEG_PTS
#define SEG_PTS 8                   // points per fine segment (arc): 8, 16 or 32
#endif
#define SEG_GROUPS (32 / SEG_PTS)   // arcs a warp looks at in one pass (lane group g takes the g-th surviving arc)
#define SEG_SHIFT (SEG_PTS == 32 ? 5 : (SEG_PTS == 16 ? 4 : 3))
#define SEG_CELLS 1024              // 16 x 4 x 16 cells of the coarse grouping

struct SegCloud {                   // one target cloud of one scan
    const float4 *pts;              // dense ring-major points; the scan id upstream's partner loops read is int(w)
    const float4 *fbox;             // [nseg][2]: lo.xyz | first dense index (int bits), hi.xyz | SEG_META(min scan id, max scan id, count) (int bits)
    const float4 *cbox;             // [ncoarse][2]: lo.xyz | member count, hi.xyz
    const float4 *mbox;             // [nseg][2]: the fine boxes again, in coarse-group order (group c = entries 32 c .. 32 c + 31): one trip, no indirection
    const int *seg_ring;            // [R + 1]
    const int *ring_start;          // [R + 1] dense index of ring r's first point
    const int4 *prange;             // [n] per target point: where upstream's backward / forward partner loops started at it break (.x, .y) and the arcs [.z, .w) in between
    int nseg, ncoarse, n_rings;
};

// Scan ids.  Upstream's partner loops read the scan id of a target point as int(intensity) (oracle/laser_odometry.c
// partners_*), NOT the ring the point was filed under: intensity = ring + relTime and relTime is slightly negative for points
// whose azimuth precedes the sweep's first point, so such a point (or a less-flat centroid whose first member is one) reads as
// ring - 1.  The index therefore classifies by int(w) per point, keeps the min / max scan id of every arc for pruning, and
// takes the loops' break positions from k3_partner_ranges.
#define SEG_EBIAS 256
#define SEG_META(emin, emax, cnt) ((min(max((emin) + SEG_EBIAS, 0), 1023) << 18) | (min(max((emax) + SEG_EBIAS, 0), 1023) << 8) | (cnt))
#define SEG_META_EMIN(meta) ((int)(((unsigned)(meta)) >> 18) - SEG_EBIAS)
#define SEG_META_EMAX(meta) ((int)((((unsigned)(meta)) >> 8) & 1023u) - SEG_EBIAS)

__device__ __forceinline__ float seg_box_lb2(const float4 lo, const float4 hi, float qx, float qy, float qz)
{
    const float gx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.0f);
    const float gy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.0f);
    const float gz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.0f);
    return (gx * gx + gy * gy) + gz * gz;
}

// candidate admission + tie rule.  mode 0: plain nearest neighbour (tie = dense index, lowest wins);
// mode 1: upstream's partner loops (oracle/laser_odometry.c partners_corner / partners_surf) around the nearest
// neighbour `ind` whose scan id is `scan`: the forward loop visits the dense indices (ind, hi), the backward loop (lo, ind)
// (lo / hi: where the loops break, k3_partner_ranges; hi also carries the forward-bound quirk); a forward point belongs to
// the same-scan class (want 2) if its scan id is <= scan, a backward point if it is >= scan, everything else to the
// other-scan class (want 3).  Ties: forward indices first (ascending), then backward (descending), as the loops visit them.
struct SegFilter {
    int mode, ind, scan, lo, hi, want;
    int f0, f1;                 // the arcs that hold the dense indices (lo, hi)  (k3_partner_ranges)
    // can an arc (dense indices [s0, s0 + cnt), scan ids within [emin, emax]) hold a point of class `cls`?
    __device__ __forceinline__ bool arc_may(int s0, int meta, int cls) const
    {
        const int cnt = meta & 0xff;
        if (s0 + cnt - 1 <= lo || s0 >= hi) return false;
        if ((unsigned)(scan + SEG_EBIAS - 1) >= 1021u) return true;                  // scan id outside what the arc summaries can express
        const int emin = SEG_META_EMIN(meta), emax = SEG_META_EMAX(meta);
        const bool fwd = s0 + cnt - 1 > ind, bwd = s0 < ind;                          // the arc holds forward / backward points
        return cls == 2 ? ((fwd && emin <= scan) || (bwd && emax >= scan)) : ((fwd && emax > scan) || (bwd && emin < scan));
    }
    // `cls`: the class asked for (the two partner searches of a surface point share one filter and differ only in it)
    __device__ __forceinline__ bool admit(int e, int idx, unsigned &tie, int cls_wanted) const
    {
        if (mode == 0) { tie = (unsigned)idx; return true; }
        if (idx <= lo || idx >= hi || idx == ind) return false;
        int cls;
        if (idx > ind) { cls = e > scan ? 3 : 2; tie = (unsigned)(idx - ind); }
        else { cls = e < scan ? 3 : 2; tie = 0x40000000u + (unsigned)(ind - idx); }
        return cls == cls_wanted;
    }
};

struct SegBest { unsigned d, t; int idx; };      // lane-local best: d = float bits of d2 (or of dmax: none yet)

// `meta`: the SEG_META of the arc that holds `idx`, or 0 (unknown).  An arc whose points all carry one scan id (min == max:
// every arc of a sweep without negative relTime) is filtered on that id, before the point is loaded; otherwise on int(w).
__device__ __forceinline__ void seg_consider(const SegCloud &c, int idx, int meta, float qx, float qy, float qz, float dmax,
                                             const SegFilter &flt, int want, SegBest &best)
{
    unsigned tie;
    float4 p;
    const int emin = SEG_META_EMIN(meta);
    if (meta != 0 && emin == SEG_META_EMAX(meta) && (unsigned)(emin + SEG_EBIAS - 1) < 1022u) {      // (not a clamped summary)
        if (!flt.admit(emin, idx, tie, want)) return;
        p = c.pts[idx];
    } else {
        p = c.pts[idx];
        if (!flt.admit((int)p.w, idx, tie, want)) return;
    }
    const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
    const float d2 = (dx * dx + dy * dy) + dz * dz;
    const unsigned db = __float_as_uint(d2);
    if (d2 < dmax && (db < best.d || (db == best.d && tie < best.t))) { best.d = db; best.t = tie; best.idx = idx; }
}

// the surviving arcs of `fmask` (bit k = the arc whose box lane k has just tested), SEG_GROUPS of them per pass: lane
// group g looks at the g-th set bit, one point per lane.  `s0_mine` / `meta_mine`: first dense index and SEG_META (count in the low byte)
// of this lane's arc, broadcast from the registers of the lane that tested it (no second trip to the box arrays).
__device__ __forceinline__ void seg_scan_mask(const SegCloud &c, unsigned fmask, int s0_mine, int meta_mine, float qx, float qy, float qz,
                                              float dmax, const SegFilter &flt, int want, SegBest &best, int lane)
{
    const int g = lane >> SEG_SHIFT, l = lane & (SEG_PTS - 1);
    while (fmask) {
        // position of the (g + 1)-th set bit of fmask (none: fewer bits than that)
        unsigned m = fmask;
        #pragma unroll
        for (int k = 0; k < SEG_GROUPS - 1; k++) if (k < g) m &= m - 1u;
        const int bit = m ? __ffs(m) - 1 : 0;
        const int s0 = __shfl_sync(0xffffffffu, s0_mine, bit), meta = __shfl_sync(0xffffffffu, meta_mine, bit);
        if (m && l < (meta & 0xff)) seg_consider(c, s0 + l, meta, qx, qy, qz, dmax, flt, want, best);
        #pragma unroll
        for (int k = 0; k < SEG_GROUPS; k++) fmask &= fmask - 1u;       // drop the SEG_GROUPS lowest set bits
    }
}

// Exact warp-cooperative search.  flt.mode 0: every point (coarse groups first); mode 1: only the arcs that overlap the
// dense-index range (flt.lo, flt.hi) of upstream's partner loops -- one contiguous range of the ring-major numbering --
// and can hold a point of class flt.want.
// seed >= 0: a point known to be a candidate (the previous association round's answer): the search starts from its
// distance, so nearly every box is pruned at once.
// Returns the dense index of the (d2, tie) minimum among admissible points with d2 < dmax, or -1; every lane gets it.
__device__ __forceinline__ int seg_search(const SegCloud &c, float qx, float qy, float qz,
                                          float dmax, const SegFilter &flt, int seed, int lane, int *ring_out)
{
    const unsigned dmaxb = __float_as_uint(dmax);
    SegBest best; best.d = dmaxb; best.t = 0xFFFFFFFFu; best.idx = -1;
    if (seed >= 0) seg_consider(c, seed, 0, qx, qy, qz, dmax, flt, flt.want, best);      // same value in every lane
    unsigned bound = best.d;                     // float bits of the best d2 any lane holds (dmax: none yet)
    if (flt.mode == 0) {
        if (best.idx < 0) {
            // ---- phase A: the coarse group nearest to the query, its nearest fine segment -> a first bound
            float my = __int_as_float(0x7f800000); int myc = -1;
            for (int cb = lane; cb < c.ncoarse; cb += 32) {
                const float lb = seg_box_lb2(c.cbox[2 * cb], c.cbox[2 * cb + 1], qx, qy, qz);
                if (lb < my) { my = lb; myc = cb; }
            }
            const unsigned m = __reduce_min_sync(0xffffffffu, myc >= 0 ? __float_as_uint(my) : 0x7f800000u);
            if (!(m < dmaxb)) { if (ring_out) *ring_out = 0; return -1; }      // nothing closer than dmax (also: empty cloud, NaN query)
            const int src = __ffs(__ballot_sync(0xffffffffu, myc >= 0 && __float_as_uint(my) == m)) - 1;
            const int cstar = __shfl_sync(0xffffffffu, myc, src);
            const int nmem = __float_as_int(c.cbox[2 * cstar].w);
            int s0 = 0, meta = 0; float lbf = __int_as_float(0x7f800000);
            if (lane < nmem) {
                const float4 lo = c.mbox[2 * (32 * cstar + lane)], hi = c.mbox[2 * (32 * cstar + lane) + 1];
                s0 = __float_as_int(lo.w); meta = __float_as_int(hi.w);
                lbf = seg_box_lb2(lo, hi, qx, qy, qz);
            }
            const unsigned mf = __reduce_min_sync(0xffffffffu, __float_as_uint(lbf));
            const int srcf = __ffs(__ballot_sync(0xffffffffu, lane < nmem && __float_as_uint(lbf) == mf)) - 1;
            seg_scan_mask(c, 1u << srcf, s0, meta, qx, qy, qz, dmax, flt, flt.want, best, lane);
            bound = __reduce_min_sync(0xffffffffu, best.d);
        }
        // ---- phase B: every coarse group / fine segment whose box can hold a point at least as close
        for (int cb0 = 0; cb0 < c.ncoarse; cb0 += 32) {
            const int cb = cb0 + lane;
            float lbc = __int_as_float(0x7f800000);
            if (cb < c.ncoarse) lbc = seg_box_lb2(c.cbox[2 * cb], c.cbox[2 * cb + 1], qx, qy, qz);
            unsigned cmask = __ballot_sync(0xffffffffu, __float_as_uint(lbc) <= bound && lbc < dmax);
            while (cmask) {
                const int j = __ffs(cmask) - 1;
                cmask &= cmask - 1u;
                if (__float_as_uint(__shfl_sync(0xffffffffu, lbc, j)) > bound) continue;      // the bound has tightened since the ballot
                const int cc = cb0 + j;
                const int nmem = __float_as_int(c.cbox[2 * cc].w);
                int s0 = 0, meta = 0; float lbf = __int_as_float(0x7f800000);
                if (lane < nmem) {
                    const float4 lo = c.mbox[2 * (32 * cc + lane)], hi = c.mbox[2 * (32 * cc + lane) + 1];
                    s0 = __float_as_int(lo.w); meta = __float_as_int(hi.w);
                    lbf = seg_box_lb2(lo, hi, qx, qy, qz);
                }
                const unsigned fmask = __ballot_sync(0xffffffffu, __float_as_uint(lbf) <= bound && lbf < dmax);
                seg_scan_mask(c, fmask, s0, meta, qx, qy, qz, dmax, flt, flt.want, best, lane);
                bound = __reduce_min_sync(0xffffffffu, best.d);
            }
        }
    } else if (flt.hi - flt.lo > 1) {
        const int f0 = flt.f0, f1 = flt.f1;
        if (best.idx < 0) {
            // ---- phase A: nearest admissible fine segment of the range -> a first bound
            float my = __int_as_float(0x7f800000); int mys0 = 0, mymeta = -1;
            for (int f = f0 + lane; f < f1; f += 32) {
                const float4 lo = c.fbox[2 * f], hi = c.fbox[2 * f + 1];
                if (!flt.arc_may(__float_as_int(lo.w), __float_as_int(hi.w), flt.want)) continue;
                const float lb = seg_box_lb2(lo, hi, qx, qy, qz);
                if (lb < my) { my = lb; mys0 = __float_as_int(lo.w); mymeta = __float_as_int(hi.w); }
            }
            const unsigned m = __reduce_min_sync(0xffffffffu, mymeta >= 0 ? __float_as_uint(my) : 0x7f800000u);
            if (!(m < dmaxb)) { if (ring_out) *ring_out = 0; return -1; }
            const int src = __ffs(__ballot_sync(0xffffffffu, mymeta >= 0 && __float_as_uint(my) == m)) - 1;
            seg_scan_mask(c, 1u << src, mys0, mymeta, qx, qy, qz, dmax, flt, flt.want, best, lane);
            bound = __reduce_min_sync(0xffffffffu, best.d);
        }
        // ---- phase B
        for (int fb0 = f0; fb0 < f1; fb0 += 32) {
            const int f = fb0 + lane;
            float lbf = __int_as_float(0x7f800000); int s0 = 0, meta = 0;
            if (f < f1) {
                const float4 lo = c.fbox[2 * f], hi = c.fbox[2 * f + 1];
                s0 = __float_as_int(lo.w); meta = __float_as_int(hi.w);
                lbf = seg_box_lb2(lo, hi, qx, qy, qz);
                if (__float_as_uint(lbf) <= bound && lbf < dmax && !flt.arc_may(s0, meta, flt.want)) lbf = __int_as_float(0x7f800000);
            }
            const unsigned fmask = __ballot_sync(0xffffffffu, __float_as_uint(lbf) <= bound && lbf < dmax);
            seg_scan_mask(c, fmask, s0, meta, qx, qy, qz, dmax, flt, flt.want, best, lane);
            bound = __reduce_min_sync(0xffffffffu, best.d);
        }
    }
    // ---- (d2, tie) minimum over the lanes
    const unsigned md = __reduce_min_sync(0xffffffffu, best.idx >= 0 ? best.d : 0xFFFFFFFFu);
    if (md == 0xFFFFFFFFu) { if (ring_out) *ring_out = 0; return -1; }
    const unsigned mt = __reduce_min_sync(0xffffffffu, (best.idx >= 0 && best.d == md) ? best.t : 0xFFFFFFFFu);
    const int src = __ffs(__ballot_sync(0xffffffffu, best.idx >= 0 && best.d == md && best.t == mt)) - 1;
    const int idx = __shfl_sync(0xffffffffu, best.idx, src);
    if (ring_out) *ring_out = (int)c.pts[idx].w;
    return idx;
}

// The two partner searches of a surface point in ONE walk over the arcs of the loops' dense-index range: arcs that can hold
// same-scan points feed the same-scan partner (class 2), arcs that can hold other-scan points the other-scan partner
// (class 3; nearly always an arc is one or the other); each class keeps its own bound.  One filter serves both: ind / scan /
// lo / hi.  Same results as two seg_search calls.
__device__ __forceinline__ void seg_search_partners(const SegCloud &c, float qx, float qy, float qz, float dmax,
                                                    const SegFilter &f2, int seed2, int seed3, int lane, int &i2, int &i3)
{
    const unsigned dmaxb = __float_as_uint(dmax);
    const float INF = __int_as_float(0x7f800000);
    SegBest b2, b3;
    b2.d = b3.d = dmaxb; b2.t = b3.t = 0xFFFFFFFFu; b2.idx = b3.idx = -1;
    if (seed2 >= 0) seg_consider(c, seed2, 0, qx, qy, qz, dmax, f2, 2, b2);
    if (seed3 >= 0) seg_consider(c, seed3, 0, qx, qy, qz, dmax, f2, 3, b3);
    int f0 = 0, f1 = 0;
    if (f2.hi - f2.lo > 1) { f0 = f2.f0; f1 = f2.f1; }
    const bool need2 = b2.idx < 0, need3 = b3.idx < 0;          // warp-uniform (the seeds are)
    if (need2 || need3) {
        // ---- phase A: the nearest arc of each class that has no seed -> first bounds
        float my2 = INF, my3 = INF; int s2 = 0, m2 = -1, s3 = 0, m3 = -1;
        for (int f = f0 + lane; f < f1; f += 32) {
            const float4 lo = c.fbox[2 * f], hi = c.fbox[2 * f + 1];
            const int s0 = __float_as_int(lo.w), meta = __float_as_int(hi.w);
            const float lb = seg_box_lb2(lo, hi, qx, qy, qz);
            if (f2.arc_may(s0, meta, 2) && lb < my2) { my2 = lb; s2 = s0; m2 = meta; }
            if (f2.arc_may(s0, meta, 3) && lb < my3) { my3 = lb; s3 = s0; m3 = meta; }
        }
        if (need2) {
            const unsigned m = __reduce_min_sync(0xffffffffu, m2 >= 0 ? __float_as_uint(my2) : 0x7f800000u);
            if (m < dmaxb) {
                const int src = __ffs(__ballot_sync(0xffffffffu, m2 >= 0 && __float_as_uint(my2) == m)) - 1;
                seg_scan_mask(c, 1u << src, s2, m2, qx, qy, qz, dmax, f2, 2, b2, lane);
            }
        }
        if (need3) {
            const unsigned m = __reduce_min_sync(0xffffffffu, m3 >= 0 ? __float_as_uint(my3) : 0x7f800000u);
            if (m < dmaxb) {
                const int src = __ffs(__ballot_sync(0xffffffffu, m3 >= 0 && __float_as_uint(my3) == m)) - 1;
                seg_scan_mask(c, 1u << src, s3, m3, qx, qy, qz, dmax, f2, 3, b3, lane);
            }
        }
    }
    unsigned bound2 = __reduce_min_sync(0xffffffffu, b2.d), bound3 = __reduce_min_sync(0xffffffffu, b3.d);
    // ---- phase B
    for (int fb0 = f0; fb0 < f1; fb0 += 32) {
        const int f = fb0 + lane;
        float lbf = INF; int s0 = 0, meta = 0; bool may2 = false, may3 = false;
        if (f < f1) {
            const float4 lo = c.fbox[2 * f], hi = c.fbox[2 * f + 1];
            s0 = __float_as_int(lo.w); meta = __float_as_int(hi.w);
            lbf = seg_box_lb2(lo, hi, qx, qy, qz);
            // the class tests only for the few arcs whose box survives a bound (nearly every arc of the range is admissible for one
            // class or the other: testing them first cost 30 instructions per arc)
            const bool c2 = lbf < dmax && __float_as_uint(lbf) <= bound2, c3 = lbf < dmax && __float_as_uint(lbf) <= bound3;
            if (c2) may2 = f2.arc_may(s0, meta, 2);
            if (c3) may3 = f2.arc_may(s0, meta, 3);
        }
        const unsigned k2 = __ballot_sync(0xffffffffu, may2);
        const unsigned k3 = __ballot_sync(0xffffffffu, may3);
        if (k2) { seg_scan_mask(c, k2, s0, meta, qx, qy, qz, dmax, f2, 2, b2, lane); bound2 = __reduce_min_sync(0xffffffffu, b2.d); }
        if (k3) { seg_scan_mask(c, k3, s0, meta, qx, qy, qz, dmax, f2, 3, b3, lane); bound3 = __reduce_min_sync(0xffffffffu, b3.d); }
    }
    // ---- (d2, tie) minima
    {
        const unsigned md = __reduce_min_sync(0xffffffffu, b2.idx >= 0 ? b2.d : 0xFFFFFFFFu);
        const unsigned mt = __reduce_min_sync(0xffffffffu, (b2.idx >= 0 && b2.d == md) ? b2.t : 0xFFFFFFFFu);
        const unsigned who = __ballot_sync(0xffffffffu, b2.idx >= 0 && b2.d == md && b2.t == mt);
        i2 = who ? __shfl_sync(0xffffffffu, b2.idx, __ffs(who) - 1) : -1;
    }
    {
        const unsigned md = __reduce_min_sync(0xffffffffu, b3.idx >= 0 ? b3.d : 0xFFFFFFFFu);
        const unsigned mt = __reduce_min_sync(0xffffffffu, (b3.idx >= 0 && b3.d == md) ? b3.t : 0xFFFFFFFFu);
        const unsigned who = __ballot_sync(0xffffffffu, b3.idx >= 0 && b3.d == md && b3.t == mt);
        i3 = who ? __shfl_sync(0xffffffffu, b3.idx, __ffs(who) - 1) : -1;
    }
}
