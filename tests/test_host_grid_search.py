"""Host emulation of the thread-level voxel-hash searches (csrc/grid.cuh compiled for the CPU, tests/host/): exactness
against a brute-force scan.  Runs without a GPU, so the search logic behind the scan-to-map (k5_assoc) and batch
scan-to-scan (k3_assoc_thread) association is checked on every CPU run as well as by the -m gpu parity tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "grid_search_host.cpp")
OUT = os.path.join(ROOT, "tests", "host", "_build", "libgridhost.so")


@pytest.fixture(scope="module")
def gh():
    inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    deps = [SRC, os.path.join(ROOT, "vil_sensor_fusion_b200", "csrc", "grid.cuh"), os.path.join(ROOT, "vil_sensor_fusion_b200", "csrc", "vlo_internal.cuh"),
            os.path.join(ROOT, "oracle", "detmath.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        # -ffp-contract=off mirrors nvcc --fmad=false: d2 = ((dx*dx) + (dy*dy)) + (dz*dz) in separate IEEE operations
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I" + inc, SRC, "-o", OUT], check=True)
    return C.CDLL(OUT)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _d2(pts, q):
    d = pts[:, :3] - q[None, :3]
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]          # float32, same operation order


def _brute_knn5(pts, q, dmax):
    idx = np.full((len(q), 5), -1, np.int32)
    d2o = np.full((len(q), 5), np.float32(dmax), np.float32)
    for k in range(len(q)):
        d2 = _d2(pts, q[k])
        ok = np.nonzero(d2 < np.float32(dmax))[0]
        if len(ok) >= 5:                                   # the kernels use the result only when all five exist
            order = ok[np.lexsort((ok, d2[ok].view(np.uint32)))][:5]
            idx[k] = order
            d2o[k] = d2[order]
    return idx, d2o


def _map_like_cloud(rng, n_side=60):
    # a floor and two walls on a 0.4 m lattice with jitter, exact duplicates and an empty region
    g = (np.arange(n_side, dtype=np.float32) * np.float32(0.4)) - np.float32(12.0)
    xx, yy = np.meshgrid(g, g)
    floor = np.stack([xx.ravel(), np.full(xx.size, -1.0, np.float32), yy.ravel()], 1)
    wall = np.stack([xx.ravel(), yy.ravel() * np.float32(0.25) + np.float32(2.0), np.full(xx.size, 7.3, np.float32)], 1)
    wall2 = np.stack([np.full(xx.size, -5.05, np.float32), yy.ravel() * np.float32(0.25) + np.float32(2.0), xx.ravel()], 1)
    pts = np.concatenate([floor, wall, wall2]).astype(np.float32)
    pts[::3] += rng.normal(0, 0.03, (len(pts[::3]), 3)).astype(np.float32)
    pts[100:160] = pts[0:60]                                # exact duplicates: ties resolved by the lower index
    return np.concatenate([pts, np.zeros((len(pts), 1), np.float32)], 1)


@pytest.mark.parametrize("mode,cell", [(0, 1.0625), (0, 1.5), (1, 1.0625), (1, 0.6), (1, 0.31)])
@pytest.mark.parametrize("bound_mode", [0, 1])
def test_thread_knn5_is_exact(gh, mode, cell, bound_mode):
    rng = np.random.default_rng(7)
    pts = _map_like_cloud(rng)
    q = pts[rng.choice(len(pts), 700, replace=False), :3].copy()
    q[:500] += rng.normal(0, 0.2, (500, 3)).astype(np.float32)          # near the surfaces
    q[500:560] = pts[0:60, :3]                                            # d2 = 0 ties between i and i + 100
    q[560:600] += np.float32(40.0)                                        # nothing within range
    q[600:650, 1] += np.float32(0.9)                                      # about one search radius away: fewer than 5 in range
    # cell-border cases: queries on exact multiples of the cell edge
    q[650:700] = (np.round(q[650:700] / np.float32(cell)) * np.float32(cell)).astype(np.float32)
    q = np.ascontiguousarray(q, np.float32)
    idx = np.zeros((len(q), 5), np.int32)
    d2 = np.zeros((len(q), 5), np.float32)
    gh.host_knn5(_p(pts), len(pts), C.c_float(cell), _p(q), len(q), C.c_float(1.0), mode, bound_mode, _p(idx), _p(d2))
    bi, bd = _brute_knn5(pts, q, 1.0)
    full = bi[:, 4] >= 0
    assert full.sum() > 400 and (~full).sum() > 30
    np.testing.assert_array_equal(idx[full], bi[full])
    np.testing.assert_array_equal(d2[full].view(np.uint32), bd[full].view(np.uint32))
    assert np.all(idx[~full][:, 4] == -1)                                 # fewer than five within range: flagged, never invented


def test_partner_search_fast_path_is_exact_or_undecided(gh):
    """k3_assoc_thread's stage search: a hit of the 27-cell search limited to one cell edge is the global filtered
    (d2, tie) minimum; where it reports 'undecided' no admissible candidate exists inside that radius."""
    rng = np.random.default_rng(11)
    n_rings, per = 16, 400
    az = np.linspace(-np.pi, np.pi, per, endpoint=False, dtype=np.float32)
    pts, ring = [], []
    for r in range(n_rings):                                              # ring-major cloud, like a less-flat cloud
        rad = np.float32(6.0) + np.float32(0.3) * rng.standard_normal(per).astype(np.float32)
        el = np.float32(np.deg2rad(-15 + 2 * r))
        pts.append(np.stack([rad * np.cos(az), rad * np.tan(el) * np.ones_like(az), rad * np.sin(az), np.zeros_like(az)], 1))
        ring.append(np.full(per, r, np.int32))
    pts = np.ascontiguousarray(np.concatenate(pts), np.float32)
    ring = np.ascontiguousarray(np.concatenate(ring), np.int32)
    n = len(pts)
    nq = 600
    src = rng.choice(n, nq, replace=False)
    q = np.ascontiguousarray(pts[src, :3] + rng.normal(0, 0.05, (nq, 3)).astype(np.float32), np.float32)
    ind = src.astype(np.int32)                                            # "nearest neighbour" = the point the query came from
    same = rng.random(nq) < 0.4                                           # same-ring partner vs adjacent-ring partner
    lo = np.where(same, ring[src], ring[src] - 2).astype(np.int32)
    hi = np.where(same, ring[src], ring[src] + 2).astype(np.int32)
    skip = np.where(same, -1, ring[src]).astype(np.int32)
    for cell, fwd in ((0.7, n), (0.35, n), (1.0, n // 2)):
        out = np.zeros(nq, np.int32)
        gh.host_partner(_p(pts), _p(ring), n, C.c_float(cell), _p(q), nq, _p(ind), _p(lo), _p(hi), _p(skip), fwd, _p(out))
        edge = np.float32(cell) - np.float32(2e-3) * np.float32(cell)
        dfast = min(np.float32(25.0), edge * edge)
        decided = 0
        for k in range(nq):
            d2 = _d2(pts, q[k])
            idx = np.arange(n)
            ok = (ring >= lo[k]) & (ring <= hi[k]) & (ring != skip[k]) & (idx != ind[k]) & ((idx < ind[k]) | (idx < fwd)) & (d2 < np.float32(25.0))
            best = -1
            if ok.any():
                c = idx[ok]
                tie = np.where(c > ind[k], c - ind[k], 0x40000000 + (ind[k] - c)).astype(np.int64)
                best = c[np.lexsort((tie, d2[c].view(np.uint32)))][0]
            if best >= 0 and d2[best] < dfast:
                assert out[k] == best, (cell, k, out[k], best)
                decided += 1
            else:
                assert out[k] == -2, (cell, k, out[k], best)
        assert decided > nq // 3


def test_detmath_product_text_equals_oracle_text(gh):
    """D1: vlo_sincosf / vlo_atanf / vlo_atan2f of csrc/vlo_internal.cuh and orc_* of oracle/detmath.h, both
    compiled without FMA contraction, agree bit for bit on two million inputs (angles, slopes, huge and tiny values,
    signed zeros, infinities, NaN)."""
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-7, 7, 600000), rng.standard_normal(400000) * 1e3, rng.standard_normal(400000) * 1e-4,
                        np.float32(10.0) ** rng.uniform(-30, 30, 400000) * rng.choice([-1, 1], 400000),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 2.414213562373095, 0.4142135623730950]]).astype(np.float32)
    y = rng.permutation(x).astype(np.float32)
    gh.host_detmath_mismatches.restype = C.c_long
    bad = gh.host_detmath_mismatches(_p(x), _p(y), C.c_long(len(x)))
    assert bad == 0, bad


def test_map_linearisation_text_equals_oracle(gh, orc):
    """csrc/map_lin.cuh (eig3_jacobi, lstsq53, map_edge_coeff, map_plane_coeff, to_map) and csrc/dense6.cuh
    (vlo_solve6_colpiv_qr) compiled for the CPU give the oracle's results bit for bit: planar / linear / scattered /
    rank-deficient neighbour sets, near and far query points."""
    L = orc.lib()
    rng = np.random.default_rng(5)
    f32 = np.float32

    def bits(a):
        return np.ascontiguousarray(a, f32).view(np.uint32)

    n_edge_kept = n_plane_kept = 0
    for trial in range(4000):
        kind = trial % 5
        c = rng.uniform(-30, 30, 3)
        if kind == 0:                                       # points on a plane (+ small noise)
            u, v = np.linalg.qr(rng.standard_normal((3, 2)))[0].T
            nb = c + np.outer(rng.uniform(-0.6, 0.6, 5), u) + np.outer(rng.uniform(-0.6, 0.6, 5), v) + rng.normal(0, 0.01, (5, 3))
        elif kind == 1:                                     # points along a line
            d = rng.standard_normal(3); d /= np.linalg.norm(d)
            nb = c + np.outer(rng.uniform(-0.6, 0.6, 5), d) + rng.normal(0, 0.005, (5, 3))
        elif kind == 2:                                     # a blob
            nb = c + rng.normal(0, 0.3, (5, 3))
        elif kind == 3:                                     # lattice plane through the origin region: rank-deficient A x = -1
            nb = np.stack([rng.integers(-2, 3, 5) * 0.4, np.zeros(5), rng.integers(-2, 3, 5) * 0.4], 1)
        else:                                               # duplicates
            nb = np.repeat(c[None] + rng.normal(0, 0.2, (1, 3)), 5, 0); nb[3:] += rng.normal(0, 0.2, (2, 3))
        nb4 = np.ascontiguousarray(np.concatenate([nb, np.zeros((5, 1))], 1), f32)
        sel = np.ascontiguousarray(np.concatenate([nb4[:, :3].mean(0) + rng.normal(0, 0.15, 3), [3.25]]), f32)
        # 3x3 eigen on the neighbours' covariance (float32 as the kernels form it is not needed here: any symmetric matrix)
        M = np.cov(nb4[:, :3].T.astype(np.float64)).astype(f32)
        M = ((M + M.T) * f32(0.5)).astype(f32)
        e1, v1, e2, v2 = np.zeros(3, f32), np.zeros(9, f32), np.zeros(3, f32), np.zeros(9, f32)
        gh.host_eig3(_p(M), _p(e1), _p(v1)); L.orc_eig3_jacobi(_p(M), _p(e2), _p(v2))
        assert np.array_equal(bits(e1), bits(e2)) and np.array_equal(bits(v1), bits(v2)), ("eig3", trial)
        A = np.ascontiguousarray(nb4[:, :3]); b = np.full(5, -1.0, f32)
        x1, x2 = np.zeros(3, f32), np.zeros(3, f32)
        gh.host_lstsq53(_p(A), _p(x1)); L.orc_lstsq53(_p(A), _p(b), _p(x2))
        assert np.array_equal(bits(x1), bits(x2)), ("lstsq53", trial, x1, x2)
        for fn_h, fn_o, tag in ((gh.host_edge_coeff, L.orc_map_edge_coeff, "edge"), (gh.host_plane_coeff, L.orc_map_plane_coeff, "plane")):
            c1, c2 = np.zeros(4, f32), np.zeros(4, f32)
            k1, k2 = fn_h(_p(sel), _p(nb4), _p(c1)), fn_o(_p(sel), _p(nb4), _p(c2))
            assert k1 == k2, (tag, trial)
            if k1:
                assert np.array_equal(bits(c1), bits(c2)), (tag, trial, c1, c2)
                if tag == "edge": n_edge_kept += 1
                else: n_plane_kept += 1
        # 6x6 normal equations from random Jacobian rows (sometimes rank-deficient)
        J = rng.standard_normal((40, 6)).astype(f32)
        if trial % 7 == 0:
            J[:, 4] = J[:, 1]
        H = np.ascontiguousarray(J.T @ J, f32); g = np.ascontiguousarray(J.T @ rng.standard_normal(40).astype(f32), f32)
        s1, s2 = np.zeros(6, f32), np.zeros(6, f32)
        gh.host_solve6(_p(H), _p(g), _p(s1)); L.orc_solve6_colpiv_qr(_p(H), _p(g), _p(s2))
        assert np.array_equal(bits(s1), bits(s2)), ("solve6", trial)
        T = np.ascontiguousarray(np.concatenate([rng.uniform(-3.2, 3.2, 3), rng.uniform(-50, 50, 3)]), f32)
        o1, o2 = np.zeros(4, f32), np.zeros((1, 4), f32)
        gh.host_to_map(_p(T), _p(sel), _p(o1)); L.orc_point_to_map(_p(T), _p(sel), 1, _p(o2))
        assert np.array_equal(bits(o1), bits(o2[0])), ("to_map", trial)
    assert n_edge_kept > 300 and n_plane_kept > 300, (n_edge_kept, n_plane_kept)


def test_odometry_linearisation_text_equals_oracle(gh, orc):
    """csrc/odom_lin.cuh (edge_coeff, plane_coeff, odom_jacobian_row) and vlo_to_start compiled for the CPU give the oracle's
    results bit for bit, before and after the robust weight switches on (iteration 5)."""
    L = orc.lib()
    rng = np.random.default_rng(9)
    f32 = np.float32

    def bits(a):
        return np.ascontiguousarray(a, f32).view(np.uint32)

    def pt(v, w=0.0):
        return np.ascontiguousarray(np.concatenate([v, [w]]), f32)

    kept = 0
    for trial in range(4000):
        c = rng.uniform(-40, 40, 3)
        a, b, t3 = (pt(c + rng.normal(0, 0.5, 3)) for _ in range(3))
        sel = pt(c + rng.normal(0, 0.2 if trial % 3 else 0.01, 3), 7.03)
        it = int(rng.integers(0, 12))
        for fn_h, fn_o, args in ((gh.host_odom_edge_coeff, L.orc_edge_coeff, (a, b)), (gh.host_odom_plane_coeff, L.orc_plane_coeff, (a, b, t3))):
            c1, c2 = np.zeros(4, f32), np.zeros(4, f32)
            k1 = fn_h(_p(sel), *[_p(x) for x in args], it, _p(c1))
            k2 = fn_o(_p(sel), *[_p(x) for x in args], it, _p(c2))
            assert k1 == k2, (trial, it)
            assert np.array_equal(bits(c1), bits(c2)), (trial, it, c1, c2)
            kept += k1
            T = np.ascontiguousarray(np.concatenate([rng.uniform(-0.3, 0.3, 3), rng.uniform(-2, 2, 3)]), f32)
            r1, r2, b1, b2 = np.zeros(6, f32), np.zeros(6, f32), np.zeros(1, f32), np.zeros(1, f32)
            gh.host_odom_jacobian_row(_p(T), _p(sel), _p(c1), _p(r1), _p(b1))
            L.orc_odom_jacobian_row(_p(T), _p(sel), _p(c2), _p(r2), _p(b2))
            assert np.array_equal(bits(r1), bits(r2)) and np.array_equal(bits(b1), bits(b2)), (trial, "jacobian")
        for deskew in (0, 1):
            cfg = orc.default_config("VLP-16", deskew=deskew)
            T = np.ascontiguousarray(np.concatenate([rng.uniform(-0.3, 0.3, 3), rng.uniform(-2, 2, 3)]), f32)
            o1 = np.zeros(4, f32)
            gh.host_to_start(_p(T), _p(sel), deskew, C.c_float(1.0 / cfg.scan_period), _p(o1))
            o2 = orc.transform_to_start(cfg, T, sel[None])[0]
            assert np.array_equal(bits(o1), bits(o2)), (trial, "to_start", deskew)
    assert kept > 3000


@pytest.fixture(scope="module")
def warp():
    inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    src = os.path.join(ROOT, "tests", "host", "warp_emul_host.cpp")
    out = os.path.join(ROOT, "tests", "host", "_build", "libwarpemul.so")
    deps = [src, os.path.join(ROOT, "vil_sensor_fusion_b200", "csrc", "dense6.cuh"), os.path.join(ROOT, "vil_sensor_fusion_b200", "csrc", "vlo_internal.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++20", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I" + inc, src, "-o", out], check=True)
    return C.CDLL(out)


def test_single_warp_jacobi_and_gn_update_equal_oracle(warp, orc):
    """csrc/dense6.cuh's warp-level code run in lock step on 32 host threads (__syncwarp = a barrier): the single-warp 6x6
    Jacobi (tournament order, three rotations per round from the same matrix) and the Gauss-Newton update with the
    degeneracy projection equal the oracle's sequential C bit for bit -- well-conditioned, degenerate (corridor-like) and
    rank-deficient normal matrices, iteration 0 (projection built) and later iterations (projection applied)."""
    L = orc.lib()
    from oracle.oracle import RegResult
    rng = np.random.default_rng(13)
    f32 = np.float32

    def bits(a):
        return np.ascontiguousarray(a, f32).view(np.uint32)

    n_deg = 0
    for trial in range(150):
        rows = rng.standard_normal((300, 6)).astype(f32) * f32(3.0)
        kind = trial % 3
        if kind == 1:
            rows[:, 3] *= f32(0.01)                        # one weak translation direction: eigenvalue below the threshold
        elif kind == 2:
            rows[:, 5] = rows[:, 4]                        # exactly rank-deficient
        bvec = (rng.standard_normal(300) * 0.05).astype(f32)
        H = np.ascontiguousarray(rows.T @ rows, f32)
        H = np.triu(H) + np.triu(H, 1).T
        e1, v1 = np.zeros(6, f32), np.zeros(36, f32)
        warp.host_eig6_warp(_p(np.ascontiguousarray(H, f32)), _p(e1), _p(v1))
        e2, v2 = orc.eig6(H)
        assert np.array_equal(bits(e1), bits(e2)) and np.array_equal(bits(v1), bits(v2.ravel())), ("eig6", trial)
        total = np.zeros(28, f32)
        total[:21] = H[np.triu_indices(6)]
        total[21:27] = (rows.T @ bvec).astype(f32)
        total[27] = f32(bvec @ bvec)
        T1 = (rng.standard_normal(6) * 0.1).astype(f32); T2 = T1.copy()
        P1 = np.eye(6, dtype=f32).ravel().copy(); deg1 = C.c_int(0); ev1 = np.zeros(6, f32); conv1 = C.c_int(0)
        res = RegResult(); conv2 = C.c_int(0)
        for it in range(3):
            tot = (total * f32(1.0 if it == 0 else 0.5 ** it)).astype(f32)      # later iterations: smaller steps
            warp.host_gn_update_warp(_p(tot), it, C.c_float(30.0), C.c_float(0.05), C.c_float(0.05), _p(T1), _p(P1), C.byref(deg1), _p(ev1), C.byref(conv1))
            L.orc_gn_update(_p(tot), it, C.c_float(30.0), C.c_float(0.05), C.c_float(0.05), _p(T2), C.byref(res), C.byref(conv2), 0)
            assert np.array_equal(bits(T1), bits(T2)), ("T", trial, it)
            assert deg1.value == res.is_degenerate and conv1.value == conv2.value, ("flags", trial, it)
            if it == 0:
                assert np.array_equal(bits(ev1), bits(np.array(res.eig[:], f32))), ("eig", trial)
                assert np.array_equal(bits(P1), bits(np.array(res.P[:], f32))), ("P", trial)
        n_deg += deg1.value
    assert 40 < n_deg < 150
