"""K0 (organise) and K1 (feature extraction) of the product executed on the CPU under the SIMT emulator of tests/host/ --
the kernels' own source and their real launchers, one host thread per CUDA thread -- against the oracle, bit for bit.
The -m gpu suite proves the same on the B200; this one runs wherever the CUDA headers are, GPU or not."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tests", "host")
BUILD = os.path.join(HOST, "_build")
CSRC = os.path.join(ROOT, "vil_sensor_fusion_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libk01emul.so")
    srcs = [os.path.join(CSRC, f) for f in ("k0_organise.cu", "k1_extract.cu", "vlo_internal.cuh")] + \
           [os.path.join(HOST, f) for f in ("k01_emul_host.cpp", "cuda_emul.h", "gen_emul.py")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        for f in ("k0_organise", "k1_extract"):
            subprocess.run([sys.executable, os.path.join(HOST, "gen_emul.py"), os.path.join(CSRC, f + ".cu"), os.path.join(BUILD, f + ".emul.cpp")], check=True)
        subprocess.run(["g++", "-O2", "-std=c++20", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I" + inc, "-I" + HOST, "-I" + BUILD,
                        "-I" + CSRC, os.path.join(HOST, "k01_emul_host.cpp"), "-o", out], check=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _run(emu, gcfg, raw):
    raw = np.ascontiguousarray(raw, np.float32)
    n, stride = raw.shape
    N = gcfg.max_points
    R = gcfg.n_rings
    cap = R * gcfg.feature_regions * 32
    o = dict(cloud=np.zeros((N, 4), np.float32), ring_start=np.zeros(R + 1, np.int32), src=np.zeros(N, np.int32),
             label=np.zeros(N, np.int8), curvature=np.zeros(N, np.float32), picked=np.zeros(N, np.uint8), counts=np.zeros(8, np.int32),
             sharp=np.zeros(cap, np.int32), lsharp=np.zeros(cap, np.int32), flat=np.zeros(cap, np.int32),
             less_flat=np.zeros((N, 4), np.float32), lsr=np.zeros(R + 1, np.int32), lfr=np.zeros(R + 1, np.int32))
    rc = emu.emu_organise_extract(C.byref(gcfg), _p(raw), n, stride, _p(o["cloud"]), _p(o["ring_start"]), _p(o["src"]), _p(o["label"]),
                                  _p(o["curvature"]), _p(o["picked"]), _p(o["counts"]), _p(o["sharp"]), _p(o["lsharp"]), _p(o["flat"]),
                                  _p(o["less_flat"]), _p(o["lsr"]), _p(o["lfr"]))
    assert rc == 0, rc
    return o


def _compare(orc, ocfg, raw, o):
    cloud_o, rs_o, src_o = orc.organise(ocfg, raw)
    nv = int(o["counts"][0])
    assert nv == len(cloud_o)
    np.testing.assert_array_equal(o["ring_start"], rs_o)
    np.testing.assert_array_equal(o["src"][:nv], src_o)
    np.testing.assert_array_equal(o["cloud"][:nv].view(np.uint32), cloud_o.view(np.uint32))
    fo = orc.extract(ocfg, cloud_o, rs_o)
    np.testing.assert_array_equal(o["curvature"][:nv].view(np.uint32), fo["curvature"].view(np.uint32))
    np.testing.assert_array_equal(o["picked"][:nv], fo["picked"])
    np.testing.assert_array_equal(o["label"][:nv], fo["label"])
    c = o["counts"]
    np.testing.assert_array_equal(o["sharp"][:c[1]], fo["sharp_idx"])
    np.testing.assert_array_equal(o["lsharp"][:c[2]], fo["less_sharp_idx"])
    np.testing.assert_array_equal(o["flat"][:c[3]], fo["flat_idx"])
    np.testing.assert_array_equal(o["lsr"], fo["less_sharp_ring_start"])
    np.testing.assert_array_equal(o["lfr"], fo["less_flat_ring_start"])
    assert c[4] == len(fo["less_flat"])
    np.testing.assert_array_equal(o["less_flat"][:c[4]].view(np.uint32), fo["less_flat"].view(np.uint32))
    return fo


@pytest.mark.parametrize("case", ["clean", "noisy_rolling", "ragged", "rotated_ring_field", "uint16_ring", "uint8_ring"])
def test_emulated_organise_and_extract_equal_oracle(emu, orc, case):
    from vil_sensor_fusion_b200 import api
    kw = {}
    if case == "clean":
        raw = scenes.vlp16_scan(0.0, rolling=False, n_az=450)
    elif case == "noisy_rolling":
        raw = scenes.vlp16_scan(0.2, noise=0.02, seed=4, rolling=True, n_az=450)
    elif case == "ragged":
        raw = scenes.ragged_scan()[::4].copy()
    else:
        base = scenes.vlp16_scan(0.1, noise=0.01, seed=2, rolling=False, n_az=450)
        c0, _, s0 = orc.organise(orc.default_config("VLP-16"), base)
        ring = np.full(len(base), -1.0, np.float32)
        ring[s0] = np.floor(c0[:, 3])
        if case == "rotated_ring_field":
            raw = np.concatenate([base, ring[:, None]], 1).astype(np.float32)
            kw = dict(rotate_input=1, input_rotation=(0.2, 0.05, -0.1), ring_field=4)
        else:
            # a Velodyne-style point: x y z intensity (float32) + an integer ring field + padding, 24 bytes; dropped points
            # carry ring 200 (out of range).  The ring order is reversed so that it cannot be confused with the angle rule.
            rec = np.zeros((len(base), 24), np.uint8)
            rec[:, :16] = base.view(np.uint8).reshape(len(base), 16)
            ids = np.where(ring >= 0, 15 - ring, 200).astype(np.uint16)
            if case == "uint16_ring":
                rec[:, 18:20] = ids.view(np.uint8).reshape(-1, 2)            # byte offset 18: not float-aligned
                kw = dict(ring_field=18, ring_field_type=1)
            else:
                rec[:, 21] = ids.astype(np.uint8)
                kw = dict(ring_field=21, ring_field_type=2)
            raw = np.ascontiguousarray(rec).view(np.float32).reshape(len(base), 6)
    ocfg = orc.default_config("VLP-16", **kw)
    gcfg = api.default_config("VLP-16", max_scans=2, max_points=8192, **kw)
    o = _run(emu, gcfg, raw)
    fo = _compare(orc, ocfg, raw, o)
    if case in ("uint16_ring", "uint8_ring"):
        nv = int(o["counts"][0])
        np.testing.assert_array_equal(np.floor(o["cloud"][:nv, 3]), 15 - ring[o["src"][:nv]])       # ids really come from the field
    if case != "ragged":
        assert len(fo["sharp_idx"]) > 20 and len(fo["flat_idx"]) > 60 and len(fo["less_flat"]) > 300


def _random_scan(rng, kind):
    """Small adversarial VLP-16-shaped clouds: what a driver can deliver, not only what a simulator does."""
    n_az = int(rng.integers(30, 260))
    az = np.sort(rng.uniform(-np.pi, np.pi, n_az)).astype(np.float32) if kind != "regular" else np.linspace(-np.pi, np.pi, n_az, endpoint=False, dtype=np.float32)
    el = np.deg2rad(np.arange(-15, 16, 2)).astype(np.float32)
    A, E = np.meshgrid(az, el, indexing="ij")                        # firing order: all 16 lasers per azimuth
    rad = (4.0 + 3.0 * np.sin(3 * A) ** 2 + rng.normal(0, 0.02, A.shape)).astype(np.float32)
    if kind == "steps":                                              # depth discontinuities -> occlusion / parallel-beam rules
        rad += (np.floor(A * 2.5) % 2).astype(np.float32) * np.float32(2.5)
    x, y, z = rad * np.cos(E) * np.cos(A), rad * np.cos(E) * np.sin(A), rad * np.sin(E)
    raw = np.stack([x, y, z, np.ones_like(x)], -1).reshape(-1, 4).astype(np.float32)
    if kind == "holes":
        keep = rng.random(len(raw)) > 0.35
        keep[:5] = True
        raw = raw[keep]
    if kind == "junk":
        idx = rng.choice(len(raw), len(raw) // 10, replace=False)
        raw[idx[: len(idx) // 3], 0] = np.nan
        raw[idx[len(idx) // 3: 2 * len(idx) // 3]] = 0.0
        raw[idx[2 * len(idx) // 3:], 2] *= 20.0                      # far outside the vertical field of view
    if kind == "few_rings":                                          # only three rings return anything; one of them is short
        ring = np.tile(np.arange(16), n_az)[: len(raw)]
        raw = raw[np.isin(ring, (2, 3, 9))]
        ring = np.tile(np.arange(16), n_az)[: 0]
        raw = raw[: max(12, len(raw) - int(rng.integers(0, 40)))]
    if kind == "duplicates":
        raw[1::2] = raw[0::2][: len(raw[1::2])]                      # zero gaps between neighbours on a ring
    if kind == "shuffled":                                           # unorganised arrival order
        raw = raw[rng.permutation(len(raw))]
    return np.ascontiguousarray(raw, np.float32)


@pytest.mark.parametrize("kind", ["regular", "steps", "holes", "junk", "few_rings", "duplicates", "shuffled"])
def test_emulated_extraction_on_adversarial_scans(emu, orc, kind):
    from vil_sensor_fusion_b200 import api
    import zlib
    rng = np.random.default_rng(zlib.crc32(kind.encode()))
    ocfg = orc.default_config("VLP-16")
    gcfg = api.default_config("VLP-16", max_scans=2, max_points=8192)
    for rep in range(2):
        raw = _random_scan(rng, kind)
        o = _run(emu, gcfg, raw)
        _compare(orc, ocfg, raw, o)
