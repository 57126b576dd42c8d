#!/bin/bash
# A/B of the K0 / K1 sub-batch sizes on the headline step.  Usage: bash tools/gpu_ab.sh tag
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for V in "-1 -1" "0 0" "6 6" "24 24" "0 -1" "-1 0"; do
  set -- $V
  VLO_K0_SUB=$1 VLO_K1_SUB=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --legs none > $OUT/bench_$1_$2.json 2> $OUT/err_$1_$2.txt
  python - <<PY
import json
d=json.load(open("$OUT/bench_$1_$2.json"))
print("K0_SUB $1 K1_SUB $2: ms/step %.3f  k0 %.4f  k1 %.4f  launches %d" % (d["ms_per_step"], d["stages"]["k0_organise"]["avg_ms"], d["stages"]["k1_extract"]["avg_ms"], d["gpu_launches"]))
PY
done
for L in lin8 a10 lin8a10; do
  VLO_LIB_PATH=vil_sensor_fusion_b200/lib/variants/libvlo_$L.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --legs none > $OUT/bench_$L.json 2> $OUT/err_$L.txt
  python - <<PY
import json
d=json.load(open("$OUT/bench_$L.json"))
print("variant $L: ms/step %.3f  k5_assoc %.4f  k5_lin %.4f" % (d["ms_per_step"], d["stages"]["k5_assoc"]["avg_ms"], d["stages"]["k5_lin"]["avg_ms"]))
PY
done
