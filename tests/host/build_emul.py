"""TEST INFRASTRUCTURE: build tests/host/_build/libvlo_emul.so -- the whole library (every .cu of csrc/, kernels and host code)
compiled for the CPU against the SIMT emulator (cuda_emul.h) and the fake runtime (fake_cudart.cpp)."""
import concurrent.futures as cf
import os
import subprocess
import sys

HOST = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HOST))
CSRC = os.path.join(ROOT, "vil_sensor_fusion_b200", "csrc")
# VLO_EMUL_BUILD_DIR / VLO_EMUL_EXTRA_FLAGS: a second build next to the default one, e.g. with -fsanitize=address (memcheck of the
# emulated kernels: LD_PRELOAD=$(gcc -print-file-name=libasan.so) pytest tests/test_host_library.py)
BUILD = os.environ.get("VLO_EMUL_BUILD_DIR", os.path.join(HOST, "_build"))
EXTRA = os.environ.get("VLO_EMUL_EXTRA_FLAGS", "").split()
sys.path.insert(0, HOST)
import gen_emul  # noqa: E402

FLAGS = ["-O2", "-g", "-std=c++20", "-ffp-contract=off", "-fPIC", "-pthread", "-w", "-DVLO_HOST_EMULATION", "-include", os.path.join(HOST, "cuda_emul.h")]


def build(force=False):
    inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        return None
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libvlo_emul.so")
    cus = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and not f.startswith("synth_"))      # libvlo_synth.so is a GPU-only bench tool
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HOST, f) for f in ("cuda_emul.h", "fake_cudart.cpp", "gen_emul.py", "build_emul.py")] + \
           [os.path.join(ROOT, "include", "vlo.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out

    def one(f):
        gen = os.path.join(BUILD, f[:-3] + ".emul.cpp")
        open(gen, "w").write(gen_emul.rewrite(open(os.path.join(CSRC, f)).read()))
        obj = gen[:-4] + ".o"
        subprocess.run(["g++"] + FLAGS + EXTRA + ["-I" + inc, "-I" + CSRC, "-c", gen, "-o", obj], check=True)
        return obj

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, cus))
    fake = os.path.join(BUILD, "fake_cudart.o")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC"] + EXTRA + ["-I" + inc, "-c", os.path.join(HOST, "fake_cudart.cpp"), "-o", fake], check=True)
    # -Bsymbolic: the library's cuda* calls must bind to its own fake runtime even when a real libcudart is already in the
    # process (torch loads one with global visibility)
    subprocess.run(["g++", "-shared", "-pthread", "-Wl,-Bsymbolic"] + EXTRA + ["-o", out] + objs + [fake], check=True)
    return out


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
