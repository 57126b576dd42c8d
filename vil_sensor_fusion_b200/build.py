"""In-tree build of libvlo.so (all CUDA kernels + the C-ABI) for sm_100a.

nvcc cross-compiles without a GPU; the resulting vil_sensor_fusion_b200/lib/libvlo.so is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# --fmad=false is load-bearing: bit-exact parity with the oracle (see DESIGN.md "Determinism")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(paths):
    hsh = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            hsh.update(f.read())
    hsh.update(" ".join(NVCC_FLAGS).encode())
    return hsh.hexdigest()


def lib_path(name="libvlo.so"):
    return os.path.join(LIBDIR, name)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "vlo.h"))
    groups = {"libvlo.so": [s for s in _sources() if not s.startswith("synth_")],
              "libvlo_synth.so": [s for s in _sources() if s.startswith("synth_")]}
    out = lib_path()
    for lib, srcs in groups.items():
        if not srcs:
            continue
        target = lib_path(lib)
        stamp = _stamp([os.path.join(CSRC, s) for s in srcs] + headers)
        stamp_file = os.path.join(OBJDIR, lib + ".stamp")
        if not force and os.path.exists(target) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            continue

        def compile_one(src):
            obj = os.path.join(OBJDIR, src[:-3] + ".o")
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
            if verbose:
                sys.stderr.write(r.stderr)
            return obj

        with cf.ThreadPoolExecutor(max_workers=8) as ex:
            objs = list(ex.map(compile_one, srcs))
        cmd = [NVCC, "-shared", "-o", target] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
        with open(stamp_file, "w") as f:
            f.write(stamp)
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
