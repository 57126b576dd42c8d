#!/bin/bash
# Quick iteration call: parity tests + bench without the CPU legs.  Usage: bash tools/gpu_quick.sh tag [extra bench args]
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout -k 10 420 python -m pytest tests -m gpu -x -q --timeout 150 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
timeout 900 python bench.py --no-cpu-baseline "$@" > $OUT/bench.json 2> $OUT/bench.err; tail -5 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.1f scans/s  e2e %.1f (pcie %s GB/s, frac %s)  ms/step %.3f  launches %d  lat p50 %s p95 %s" % (d["value"], d["e2e"]["value"], d["e2e"].get("pcie_h2d_gbs_measured"), d["e2e"].get("frac_of_pcie"), d["ms_per_step"], d["gpu_launches"], d["latency"]["p50_ms_per_scan"], d["latency"]["p95_ms_per_scan"]))
for k,v in d["stages"].items(): print("  %-12s total %8.3f ms  avg %8.4f ms  %7.1f GB/s  frac %.4f" % (k, v["ms_total"], v["avg_ms"], v["achieved_gbs"], v["frac"]))
print("  iters", d["config"]["mean_gn_iterations"], "corr", d["mean_corr"], "ok", d["ok_registrations"])
PY
