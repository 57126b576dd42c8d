#!/bin/bash
# ncu launch list + full captures of the named kernels.  Usage: bash tools/gpu_prof.sh tag kernel1 kernel2 ...
TAG=${1:-p}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency > $OUT/bench_under_ncu.log 2>&1
python tools/launch_shares.py $OUT/launches.csv | tee $OUT/launch_shares.txt
for K in "$@"; do
  timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:^$K -s 2 -c 1 -f -o $OUT/full_$K \
      python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency > $OUT/ncu_$K.log 2>&1
  ls -la $OUT/full_$K.ncu-rep
done
