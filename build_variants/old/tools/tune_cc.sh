#!/bin/bash
for cfg in "chains=1,coop=120000" "chains=2,coop=60000" "chains=2,coop=30000" "chains=3,coop=40000" "chains=3,coop=20000" "chains=4,coop=30000" "chains=4,coop=15000" "chains=8,coop=15000"; do
  echo -n "$cfg: "
  CGRT_TUNE="$cfg" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('e2e ms', round(j['e2e']['ms_per_step'],3))
    elif 'rror' in l: print(l[-200:])
"
done
