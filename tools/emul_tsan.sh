#!/bin/bash
# Racecheck without a GPU: the CPU-emulated library (tests/host/) built with ThreadSanitizer.  Every CUDA thread of a CTA is a
# host thread, __syncthreads / __syncwarp / the warp collectives are barriers and atomics are atomics, so a shared- or
# global-memory access pair that the kernel does not order shows up as a data race.  (CTAs run one after the other: races
# BETWEEN CTAs are out of its reach.)  Usage: bash tools/emul_tsan.sh [pytest args]   (default: tests/test_host_library.py)
export VLO_EMUL_BUILD_DIR=${VLO_EMUL_BUILD_DIR:-/tmp/vlo_emul_tsan}
export VLO_EMUL_EXTRA_FLAGS="-DEMU_THREADS -fsanitize=thread -fno-omit-frame-pointer"
python tests/host/build_emul.py || exit 1
if [ $# -eq 0 ]; then set -- tests/test_host_library.py; fi
TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=2" LD_PRELOAD=$(gcc -print-file-name=libtsan.so) \
    python -m pytest "$@" -q -s -p no:cacheprovider 2>&1 | tee /tmp/vlo_emul_tsan.log | grep -E "WARNING: ThreadSanitizer|passed|failed"
echo "ThreadSanitizer warnings: $(grep -c 'WARNING: ThreadSanitizer' /tmp/vlo_emul_tsan.log)"
