"""CPU-side checks of the drop-in boundary: libvlo.so loads, exports every symbol include/vlo.h
declares, struct layouts match the header, host helpers work, and the product path fails loudly
without a GPU instead of falling back to the CPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from vil_sensor_fusion_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.load()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vlo.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vlo_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from vil_sensor_fusion_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libvlo.so does not export %s" % n
        assert n in _lib.SYMBOLS, "%s is declared in vlo.h but not bound in _lib.SYMBOLS" % n
    for n in _lib.SYMBOLS:
        assert n in names, "%s is bound but not declared in include/vlo.h" % n


def test_struct_layouts_match_header(tmp_path):
    from vil_sensor_fusion_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vlo.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(vlo_config),sizeof(vlo_result),sizeof(vlo_feature_counts),sizeof(vlo_preint),'
                   'offsetof(vlo_config,cov_accel),offsetof(vlo_result,cov),offsetof(vlo_preint,cov));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    exp = [C.sizeof(_lib.Config), C.sizeof(_lib.Result), C.sizeof(_lib.FeatureCounts), C.sizeof(_lib.Preint),
           _lib.Config.cov_accel.offset, _lib.Result.cov.offset, _lib.Preint.cov.offset]
    assert got == exp


def test_header_is_plain_c_and_cites_reference():
    hdr = open(os.path.join(ROOT, "include", "vlo.h")).read()
    assert "torch" not in hdr.lower() and 'extern "C"' in hdr
    for cite in ("degerate_odometry_filter.cpp", "IMUManager.cpp:27-74", "SensorManagerRos.cpp:122-158", "loam_params.yaml"):
        assert cite in hdr


def test_default_config_mirrors_reference_yaml(lib):
    from vil_sensor_fusion_b200 import api
    c = api.default_config("HDL-64E")
    assert (c.n_rings, c.feature_regions, c.curvature_region, c.max_corner_sharp, c.max_corner_less_sharp, c.max_surface_flat) == (64, 6, 5, 2, 20, 4)
    assert (c.odom_max_iterations, c.map_max_iterations, c.odom_degen_eig, c.map_degen_eig) == (25, 10, 30.0, 40.0)
    assert abs(c.dopt_rot_threshold - 11.5) < 1e-6 and abs(c.dopt_trans_threshold - 28.9) < 1e-5
    assert (c.cov_accel, c.cov_bias_acc, c.cov_integration) == (1e-6, 1e-4, 1e-8)
    with pytest.raises(ValueError):
        api.default_config("no-such-lidar")


def test_host_helpers_without_gpu(lib, orc):
    from vil_sensor_fusion_b200 import api
    np.testing.assert_allclose(api.pose_diff([0, 0, 0, 1, 0, 0, 0], [1, 1, 1, 1, 0, 0, 0]), [1, 1, 1, 1, 0, 0, 0])   # UnitTests.cpp:228-233
    rng = np.random.default_rng(0)
    A = rng.normal(size=(50, 6)) * np.array([40, 40, 40, 4, 4, 4])
    H = (A.T @ A).astype(np.float32)
    ok, lr, lt = api.dopt_gate(H)
    oko, lro, lto = orc.dopt_gate(H)
    assert ok == oko
    np.testing.assert_allclose([lr, lt], [lro, lto], rtol=1e-6)
    st = api.hessian_stack(np.zeros(7, api.RESULT_DTYPE))
    assert st.shape == (6, 6, 7) and st.dtype == np.float64        # make_prettier_graphs.py:440 layout


def test_no_cpu_fallback_without_device(lib):
    """On a box without a GPU the product refuses to run (VLO_ERR_NO_DEVICE); it never routes to the oracle."""
    import torch
    from vil_sensor_fusion_b200 import api
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.VloError) as e:
        api.Handle(api.default_config("VLP-16"))
    assert e.value.code == -4


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vil_sensor_fusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt and "vlo_oracle.h" not in txt, f


def test_ros_shim_compiles_against_mock_headers_and_links(tmp_path):
    """the `loam`-compatible node (integration/ros/vlo_loam_node.cpp, the code INTEGRATION.md shows) is
    compiled against the mock ROS headers under integration/mock and linked with libvlo.so, so every ABI call it makes
    exists with that signature."""
    import re
    import subprocess
    from vil_sensor_fusion_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "integration", "ros", "vlo_loam_node.cpp")
    obj = str(tmp_path / "vlo_loam_node.o")
    exe = str(tmp_path / "vlo_loam_node")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror=implicit-function-declaration", "-I", os.path.join(root, "integration", "mock"),
                        "-I", os.path.join(root, "include"), "-c", src, "-o", obj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    libdir = os.path.dirname(build.lib_path())
    r = subprocess.run(["g++", "-o", exe, obj, "-L", libdir, "-lvlo", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the document shows the same code
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```cpp\n(// vlo_loam_node\.cpp.*?)```", doc, re.S).group(1)
    assert block.strip() in open(src).read()


def test_undistort_input_cloud_is_refused_not_remapped():
    """the fork's undistortInputCloud (ego-motion compensation of the input cloud from an external prior / IMU) is not
    in the reference and not implemented here; asking for it is an explicit VLO_ERR_UNSUPPORTED at vlo_create (no GPU needed
    to find out), never a silent mapping onto LaserOdometry's `deskew`."""
    from vil_sensor_fusion_b200 import _lib
    lib = _lib.load()
    c = _lib.Config()
    lib.vlo_default_config(C.byref(c))
    assert c.undistort_input_cloud == 0 and c.deskew == 1
    c.undistort_input_cloud = 1
    h = C.c_void_p()
    assert lib.vlo_create(C.byref(c), C.byref(h)) == -6 and not h.value


def _build_imu_seam(tmp_path, emulated=False):
    import subprocess
    from vil_sensor_fusion_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = ["-I", os.path.join(root, "integration", "mock"), "-I", os.path.join(root, "include")]
    objs = []
    for name in ("IMUManager_vlo", "imu_seam_check"):
        obj = str(tmp_path / (name + ".o"))
        r = subprocess.run(["g++", "-std=c++17", "-Wall"] + inc + ["-c", os.path.join(root, "integration", "ros", name + ".cpp"), "-o", obj],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        objs.append(obj)
    exe = str(tmp_path / "imu_seam_check")
    libdir, lib = os.path.dirname(build.lib_path()), "-lvlo"
    if emulated:                                  # pytest --emulated: the CPU-emulated build of the same library (tests/host)
        libdir, lib = os.path.join(root, "tests", "host", "_build"), "-lvlo_emul"
    r = subprocess.run(["g++", "-o", exe] + objs + ["-L", libdir, lib, "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_imu_seam_replaces_imumanager_cpp_at_link_time(tmp_path):
    """SURVEY 8f N1 / VERDICT r1: GraphManager calls the NON-virtual IMUManager::getFactor through a shared_ptr<ImuManagerRos>
    (GraphManager.h:101, IMUManager.h:26-27), so the seam that binds is a translation unit compiled instead of
    IMUManager.cpp.  integration/ros/IMUManager_vlo.cpp defines every member IMUManager.cpp defines, against the
    reference's class declaration, and links with a caller that only sees that header."""
    import subprocess
    exe = _build_imu_seam(tmp_path)
    assert subprocess.run([exe]).returncode == 0                      # no GPU: link + load check only
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_hdr = "/root/reference/gtsam_fusion/include/gtsam_fusion/IMUManager.h"
    if os.path.exists(ref_hdr):                                       # (this container only) the restated declaration is the reference's
        import re
        norm = lambda s: re.sub(r"\s+", " ", re.sub(r"//.*", "", s))
        ref = norm(open(ref_hdr).read())
        mock = norm(open(os.path.join(root, "integration", "mock", "gtsam_fusion", "IMUManager.h")).read())
        for decl in ("explicit IMUManager(boost::shared_ptr<PreintegratedCombinedMeasurements::Params> imuParams);",
                     "void addIMUMeasurement(double time, const Vector3 accel, const Vector3 gyro);",
                     "CombinedImuFactor getFactor(double startTime, double endTime, uint64_t currentIndex, imuBias::ConstantBias bias);",
                     "CombinedImuFactor getFactor(double endTime, uint64_t currentIndex, imuBias::ConstantBias bias);",
                     "std::deque<Measurement> _buffer;", "PreintegratedCombinedMeasurements _integrator;"):
            assert decl in ref and decl in mock, decl


@pytest.mark.gpu
def test_imu_seam_known_answer_on_the_gpu(tmp_path, request):
    """The same executable on the B200: the reference's known-answer case (UnitTests.cpp:30-66) through the replaced
    IMUManager::getFactor -- dt 0.15, dv 0.0175, dp 0.0011875."""
    import subprocess
    exe = _build_imu_seam(tmp_path, emulated=request.config.getoption("--emulated"))
    r = subprocess.run([exe, "run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    dt, dv, dp = [float(x) for x in r.stdout.split()]
    assert abs(dt - 0.15) < 1e-12
    np.testing.assert_allclose([dv, dp], [0.0175, 0.0011875], rtol=1e-6)


@pytest.mark.parametrize("compression", ["none", "bz2", "lz4"])
def test_rosbag_v2_round_trip(tmp_path, compression):
    """the ROS-free bag reader returns exactly the PointCloud2 / Imu messages a bag holds (bags written
    by the same module per the public v2.0 format: chunks with connection + message records, none / bz2 / lz4)."""
    from vil_sensor_fusion_b200 import rosbag_io as rb
    rng = np.random.default_rng(1)
    clouds = [rng.normal(0, 10, (500 + 37 * k, 5)).astype(np.float32) for k in range(5)]
    msgs = []
    for k, c in enumerate(clouds):
        msgs.append(("/lidar", "sensor_msgs/PointCloud2", rb.POINTCLOUD2_MD5, 10.0 + 0.1 * k,
                     rb.make_pointcloud2(c, 10.0 + 0.1 * k, field_names=("x", "y", "z", "intensity", "ring"), seq=k)))
        for j in range(20):
            t = 10.0 + 0.1 * k + 0.005 * j
            msgs.append(("/imu/data", "sensor_msgs/Imu", rb.IMU_MD5, t, rb.make_imu(t, [0.1 * j, 0.2, -0.3], [0.0, 9.81, 0.01 * k])))
    path = str(tmp_path / "t.bag")
    rb.write_bag(path, msgs, compression=compression, chunk_messages=16)
    assert open(path, "rb").read(13) == b"#ROSBAG V2.0\n"
    got = list(rb.read_messages(path))
    assert len(got) == len(msgs)
    assert [g[0] for g in got] == [m[0] for m in msgs] and [g[1] for g in got] == [m[1] for m in msgs]
    np.testing.assert_allclose([g[2] for g in got], [m[3] for m in msgs], atol=1e-9)
    assert all(g[3] == m[4] for g, m in zip(got, msgs))
    loaded, imu = rb.load_bag(path, "/lidar", "/imu/data")
    assert len(loaded) == 5 and imu["t"].shape == (100,)
    for up, c in zip(loaded, clouds):
        assert up["point_step"] == 20 and up["fields"] == dict(x=0, y=4, z=8)
        np.testing.assert_array_equal(np.frombuffer(up["data"].tobytes(), np.float32).reshape(-1, 5), c)
    np.testing.assert_allclose(imu["gyro"][21], [0.1, 0.2, -0.3])
    np.testing.assert_allclose(imu["accel"][-1], [0.0, 9.81, 0.04])
    only = list(rb.read_messages(path, {"/lidar"}))
    assert len(only) == 5
    with pytest.raises(RuntimeError):
        open(str(tmp_path / "x.bag"), "wb").write(b"not a bag")
        list(rb.read_messages(str(tmp_path / "x.bag")))


def test_bench_has_no_collective_inside_rank0_only_blocks():
    """bench.py under torchrun: a collective (dist.*, the bench's barrier(), bag.gather_results) reached by rank 0 alone
    deadlocks the job (seen on 2 GPUs: rank 0 in barrier() inside the whole-bag leg, rank 1 in the final all_reduce)."""
    import ast
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    tree = ast.parse(src)

    def mentions_rank0(test):
        for node in ast.walk(test):
            if isinstance(node, ast.Compare) and isinstance(node.left, ast.Name) and node.left.id == "rank" \
                    and len(node.comparators) == 1 and isinstance(node.comparators[0], ast.Constant) and node.comparators[0].value == 0 \
                    and isinstance(node.ops[0], ast.Eq):
                return True
        return False

    bad = []
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and mentions_rank0(node.test):
            for stmt in node.body:
                for sub in ast.walk(stmt):
                    if isinstance(sub, ast.Call):
                        f = sub.func
                        if isinstance(f, ast.Name) and f.id == "barrier":
                            bad.append(("barrier", sub.lineno))
                        if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "dist":
                            bad.append(("dist." + f.attr, sub.lineno))
                        if isinstance(f, ast.Attribute) and f.attr == "gather_results":
                            bad.append(("gather_results", sub.lineno))
    assert not bad, bad


def test_lz4_frame_codec():
    """The written-out LZ4 frame codec of rosbag_io (no lz4 module in this image): a hand-assembled frame per the published
    format (linked blocks, a stored block, an overlapping match, long literal / match length encodings, content size and
    checksum fields present) decodes to the known text, and compress -> decompress round-trips compressible, random and
    empty inputs."""
    import struct
    from vil_sensor_fusion_b200 import rosbag_io as rb
    # block 1: 5 literals "abcde", then match offset 5 length 4+15+3 = 22 (overlapping repeat of "abcde"), last literals "XYZ12"
    b1 = bytes([0x5F]) + b"abcde" + struct.pack("<H", 5) + bytes([3]) + bytes([0x50]) + b"XYZ12"
    # block 2 (linked): match reaching back into block 1: 1 literal "Q", offset 33 (start of the frame), length 4, then 5 literals
    b2 = bytes([0x10]) + b"Q" + struct.pack("<H", 33) + bytes([0x50]) + b"_end_"
    stored = b"RAW!"
    frame = struct.pack("<I", rb.LZ4_MAGIC) + bytes([0x4C, 0x40]) + struct.pack("<Q", 0) + bytes([0x00])   # linked, content size + checksum
    frame += struct.pack("<I", len(b1)) + b1 + struct.pack("<I", len(b2)) + b2
    frame += struct.pack("<I", len(stored) | 0x80000000) + stored + struct.pack("<I", 0) + struct.pack("<I", 0xDEADBEEF)
    text = rb.lz4_frame_decompress(frame)
    exp1 = b"abcde" + (b"abcde" * 5)[:22] + b"XYZ12"
    assert text == exp1 + b"Q" + exp1[:4] + b"_end_" + stored, text
    rng = np.random.default_rng(0)
    cases = [b"", b"x", b"hello world " * 1000, bytes(rng.integers(0, 256, 70000, dtype=np.uint8)),
             np.tile(rng.normal(0, 1, 300).astype(np.float32), 200).tobytes(), bytes(300000)]
    for c in cases:
        enc = rb.lz4_frame_compress(c)
        assert rb.lz4_frame_decompress(enc) == c
    assert len(rb.lz4_frame_compress(cases[2])) < len(cases[2]) // 10
    assert rb.lz4_frame_decompress(rb.lz4_frame_compress(b"ab" * 50) + rb.lz4_frame_compress(b"cd" * 50)) == b"ab" * 50 + b"cd" * 50
