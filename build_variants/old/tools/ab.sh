#!/bin/bash
# A/B: product library vs every build_variants/lib_*.so (speed only), short bench each
run() { python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/frame', round(j['ms_per_step'],3), j['config']['kernel_ms_per_frame_rank0'])
    elif 'rror' in l: print(l[-200:])
"; }
echo -n "product: "; run
for f in build_variants/lib_*.so; do
  case $f in *instr*) continue;; esac
  echo -n "$(basename $f): "; CGRT_LIB=$PWD/$f run
done
