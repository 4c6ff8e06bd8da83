TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${1:-2} --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/multi_check.py 2>&1 | grep "multi_check" | tail -4
for m in p2p; do
CGRT_EXCHANGE=$m timeout 300 $TR bench.py --gpus ${1:-2} --steps 20 --warmup 3 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('$m', 'Mrays/s', round(j['value'],1), 'ms/frame', round(j['ms_per_step'],3), j['config']['kernel_ms_per_frame_rank0'], 'e2e', round(j['e2e']['value'],1), j['config']['exchange'][:30], j['config']['exchange_fallback_reason'], j['config']['handoff_timeouts'])
"
done
