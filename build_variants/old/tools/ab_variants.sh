#!/bin/bash
# time every library variant under build_variants/ (A/B builds made with CGRT_NVCC_EXTRA; speed only)
for f in build_variants/lib_*.so; do
  echo -n "$(basename $f): "
  CGRT_LIB=$PWD/$f python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/frame', round(j['ms_per_step'],3), j['config']['kernel_ms_per_frame_rank0'])
    elif 'rror' in l: print(l[-200:])
"
done
