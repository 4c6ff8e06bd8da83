#!/bin/bash
# persistent CTAs per SM sweep (CGRT_TUNE), short bench each
for b in ${@:-2 3 4 5 6 8}; do
  echo -n "blocks=$b: "
  CGRT_TUNE="blocks=$b" python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/frame', round(j['ms_per_step'],3), j['config']['kernel_ms_per_frame_rank0'])
    elif 'rror' in l: print(l[-200:])
"
done
