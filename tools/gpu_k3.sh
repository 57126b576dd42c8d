#!/bin/bash
# k3 iteration call: odometry / online / bag parity tests, stage times of the 127-pair step, one full ncu capture of k3_assoc.
# Usage: bash tools/gpu_k3.sh tag
TAG=${1:-k3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout -k 10 400 python -m pytest tests/test_gpu_odometry.py tests/test_gpu_online.py tests/test_gpu_bag.py -m gpu -x -q --timeout 150 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
timeout 200 python tools/pairs_time.py 2>&1 | tail -3 | tee $OUT/pairs_time.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k3_assoc -s 6 -c 1 -f -o $OUT/full_k3_assoc python tools/pairs_time.py > $OUT/ncu.log 2>&1
ls -la $OUT
