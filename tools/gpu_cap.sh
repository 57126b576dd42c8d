#!/bin/bash
# ncu full capture (with source) of the named kernels only.  Usage: bash tools/gpu_cap.sh tag kernel1 kernel2 ...
TAG=${1:-c}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for K in "$@"; do
  timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:^$K -s 2 -c 1 -f -o $OUT/full_$K \
      python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency > $OUT/ncu_$K.log 2>&1
  ls -la $OUT/full_$K.ncu-rep
done
