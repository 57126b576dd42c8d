for C in 0.75 0.53125 0.36; do echo "== cell $C"; VLO_MAP_CELL=$C python bench.py --no-cpu-baseline --no-latency --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f ms/step %.3f' % (d['value'], d['ms_per_step']))
for k,v in d['stages'].items(): print('  %-12s avg %8.4f ms' % (k, v['avg_ms']))
print(d['mean_corr'], d['config']['mean_gn_iterations'])
"; done
for K in k5_knn k5_lin; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K\$ -s 2 -c 1 -f -o gpurun_out/q2/full_$K python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --no-latency > gpurun_out/q2/ncu_$K.log 2>&1
done
ls -la gpurun_out/q2
