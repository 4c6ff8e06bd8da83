# usage: tools/run2.sh <nGPUs> [check]
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$2" = "check" ]; then timeout 400 $TR tools/multi_check.py 2>&1 | grep "multi_check" | tail -4; fi
CGRT_EXCHANGE=p2p timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/bench_n$N.json | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('N=$N', 'Mrays/s', round(j['value'],1), 'ms/frame', round(j['ms_per_step'],3), j['config']['kernel_ms_per_frame_rank0'], 'e2e', round(j['e2e']['value'],1), j['config']['exchange'][:30], j['config']['exchange_fallback_reason'], j['config']['handoff_timeouts'])
"
