for O in 4 5 6 8; do for F in 1480 2960 5920; do echo "== occ $O fill $F"; VLO_AL_OCC=$O VLO_AL_FILL=$F python bench.py --no-cpu-baseline --no-latency --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f ms/step %.3f  assoc_lin %.4f' % (d['value'], d['ms_per_step'], d['stages']['k5_assoc_lin']['avg_ms']))
"; done; done
