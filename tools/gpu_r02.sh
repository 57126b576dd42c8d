#!/bin/bash
# Round-2 GPU call: parity tests, smoke, bench (all legs), launch lists, full ncu captures of the named kernels.
# Usage (repo root on the GPU box): bash tools/gpu_r02.sh tag "bench kernels" "pairs kernels" [skip-tests]
TAG=${1:-r02a}; BK=${2:-}; PK=${3:-}; SKIP=${4:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP" ]; then
echo "== pytest -m gpu"
timeout -k 10 600 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
fi
echo "== bench"
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 1500 $OUT/bench.json; tail -5 $OUT/bench.err
python tools/bench_brief.py $OUT/bench.json | tee $OUT/bench_brief.txt
echo "== ncu launch list (headline step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --legs none > $OUT/bench_under_ncu.log 2>&1
python tools/launch_shares.py $OUT/launches.csv | tee $OUT/launch_shares.txt
echo "== ncu launch list (whole-bag pairs step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_pairs.csv \
    python tools/pairs_time.py > $OUT/pairs_under_ncu.log 2>&1
python tools/launch_shares.py $OUT/launches_pairs.csv | tee $OUT/launch_shares_pairs.txt
for K in $BK; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o $OUT/full_$K \
      python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency --legs none > $OUT/ncu_$K.log 2>&1
  ls -la $OUT/full_$K.ncu-rep
done
for K in $PK; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o $OUT/full_$K \
      python tools/pairs_time.py > $OUT/ncu_$K.log 2>&1
  ls -la $OUT/full_$K.ncu-rep
done
echo "== ncu: batched IMU preintegration (C4)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k6_imu -s 1 -c 1 -f -o $OUT/full_k6_imu \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency --legs c4 > $OUT/ncu_k6_imu.log 2>&1
ls -la $OUT/full_k6_imu.ncu-rep
echo "== reference arm"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-300 $OUT/bench_ref.json
echo "== ncu: map maintenance kernels of the online tick (maintained map)"
for K in k7_ins_probe k7_ins_assign k7_ins_accum k7_ins_final; do
  MAINTAINED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o $OUT/full_$K \
      python tools/online_time.py > $OUT/ncu_$K.log 2>&1
  ls -la $OUT/full_$K.ncu-rep
done
