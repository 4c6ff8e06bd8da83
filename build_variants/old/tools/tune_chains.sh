#!/bin/bash
for c in ${@:-1 2 3 4 6}; do
  echo -n "chains=$c: "
  CGRT_TUNE="chains=$c" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/frame(profiled,1 chain)', round(j['ms_per_step'],3), 'e2e ms', round(j['e2e']['ms_per_step'],3))
    elif 'rror' in l: print(l[-200:])
"
done
