#!/bin/bash
for cfg in "6 8" "4 8" "3 8" "2 8" "4 5" "2 4" "1 4"; do
  set -- $cfg
  echo -n "subleaf=$1 minleaf=$2: "
  CGRT_SUBLEAF=$1 CGRT_MINLEAF=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/frame', round(j['ms_per_step'],3), j['config']['kernel_ms_per_frame_rank0'], 'e2e ms', round(j['e2e']['ms_per_step'],3))
    elif 'rror' in l: print(l[-200:])
"
done
