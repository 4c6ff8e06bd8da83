#!/bin/bash
# fast-tree flavour sweep (speed only): reference-collapse vs binned SAH with different leaf sizes
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/frame', round(j['ms_per_step'],3), j['config']['kernel_ms_per_frame_rank0'])
    elif 'rror' in l: print(l[-200:])
"; }
echo -n "ref: "; CGRT_FAST_TREE=ref run
for l in ${@:-4 6 8}; do echo -n "sah leaf=$l: "; CGRT_SAH_LEAF=$l run; done
