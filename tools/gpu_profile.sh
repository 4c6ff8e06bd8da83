#!/bin/bash
# Full evidence pass on one B200 (run under gpurun): parity tests, contract bench (both arms), ncu launch list, ncu --set full
# usage: tools/gpu_profile.sh <tag>
tag=${1:-r01x}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_paths_fast|k_shadow_fast|k_shade_paths' --launch-skip 9 --launch-count 3 \
    -o gpurun_out/prof_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
echo "full rc=$?"
head -c 600 gpurun_out/bench_$tag.json; echo; head -c 400 gpurun_out/bench_ref_$tag.json
