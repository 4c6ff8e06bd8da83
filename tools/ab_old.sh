run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/frame', round(j['ms_per_step'],3), j['config']['kernel_ms_per_frame_rank0'])
    elif 'rror' in l: print(l[-200:])
"; }
for i in 1 2; do
echo -n "HEAD: "; run
echo -n "old a4bbef2: "; (cd build_variants/old && run)
done
