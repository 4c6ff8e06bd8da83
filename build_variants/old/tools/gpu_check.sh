#!/bin/bash
# GPU regression loop: parity tests then a short bench summary (run under gpurun)
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps ${1:-10} --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('Mrays/s', round(j['value'],1), 'ms/frame', round(j['ms_per_step'],3), j['config']['kernel_ms_per_frame_rank0'], 'e2e', round(j['e2e']['value'],1), 'roofline', j['roofline']['kernel'], round(j['roofline']['frac'],3), 'frame frac', round(j['config']['frame_roofline']['frac_of_hbm_peak'],3))
    else: print(l[-300:])
"
