#!/bin/bash
# One gpurun call: parity tests, smoke, bench (both arms), ncu launch list + full captures of the top kernels.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout -k 10 420 python -m pytest tests -m gpu -x -q --timeout 150 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency > $OUT/bench_under_ncu.log 2>&1
python tools/launch_shares.py $OUT/launches.csv | tee $OUT/launch_shares.txt
echo "== ncu full captures"
for K in k5_assoc k5_lin k1_extract k1c_lessflat k0_scatter; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K -s 2 -c 1 -f -o $OUT/full_$K \
      python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency > $OUT/ncu_$K.log 2>&1
  ls -la $OUT/full_$K.ncu-rep
done
echo "== whole-bag leg: full capture of the batch scan-to-scan association"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^k3_assoc_thread -s 1 -c 1 -f -o $OUT/full_k3_assoc_thread \
    python tools/pairs_profile.py > $OUT/ncu_k3_assoc_thread.log 2>&1
ls -la $OUT/full_k3_assoc_thread.ncu-rep
