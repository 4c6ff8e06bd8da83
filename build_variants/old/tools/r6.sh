python -m pytest tests -x -q -m gpu 2>&1 | tail -1
for b in 0 16 24 32 48; do
  echo -n "budget=$b: "
  CGRT_TUNE="budget=$b" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('N=1 ms/frame', round(j['ms_per_step'],3), 'e2e ms', round(j['e2e']['ms_per_step'],3), end=' | ')
    elif 'rror' in l: print(l[-200:])
"
  CGRT_TUNE="budget=$b,coop=15000" python tools/rank_sim.py 8 10 | cut -c1-60
done
