#!/bin/bash
# Memcheck without a GPU: the CPU-emulated library (tests/host/) built with AddressSanitizer, the emulation tests run on it.
# "Device" memory is calloc'd host memory and `__shared__` arrays are statics, so an out-of-bounds access of a kernel is a
# heap / global redzone hit.  Usage: bash tools/emul_asan.sh [pytest args]   (default: tests/test_host_library.py)
export VLO_EMUL_BUILD_DIR=${VLO_EMUL_BUILD_DIR:-/tmp/vlo_emul_asan${EMUL_THREADS:+_threads}}
# fiber mode of the emulator (fast: the whole -m gpu suite in ~8 min); EMUL_THREADS=1 for the host-thread mode
export VLO_EMUL_EXTRA_FLAGS="${EMUL_THREADS:+-DEMU_THREADS }-fsanitize=address -fno-omit-frame-pointer"
python tests/host/build_emul.py || exit 1
if [ $# -eq 0 ]; then set -- tests/test_host_library.py; fi
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:detect_stack_use_after_return=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
    python -m pytest "$@" -x -q -p no:cacheprovider
