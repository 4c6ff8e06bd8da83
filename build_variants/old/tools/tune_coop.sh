#!/bin/bash
for c in ${@:-0 60000 120000 200000 350000}; do
  echo -n "coop=$c: "
  CGRT_TUNE="coop=$c" python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/frame', round(j['ms_per_step'],3), j['config']['kernel_ms_per_frame_rank0'])
    elif 'rror' in l: print(l[-200:])
"
done
