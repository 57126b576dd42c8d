"""The GPU parity suite itself, without a GPU: `pytest tests -m gpu --emulated` as one test of the CPU suite.  Every
`-m gpu` test (full-size VLP-16 / HDL-64 scans, the 1M-point map, the online chains, the whole-bag API) runs against
tests/host/_build/libvlo_emul.so -- the library's kernels and host code compiled for the CPU against the SIMT emulator -- and
must pass exactly as it does on the B200.  About three minutes on eight cores.  TEST INFRASTRUCTURE."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_parity_suite_against_the_emulated_library():
    inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "--emulated", "-q", "-x", "-p", "no:cacheprovider", "--timeout", "900"],
                       cwd=ROOT, capture_output=True, text=True, timeout=2400)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
