#!/bin/bash
# Full evidence pass on one B200 (run under gpurun): parity tests, contract bench (both arms), ncu launch list, ncu --set full
# usage: tools/gpu_profile.sh <tag>   (outputs under gpurun_out/, kept below the 64 MiB pull limit: raw CSV pages, small reps)
tag=${1:-r01x}
mkdir -p gpurun_out
if [ -z "$QUICK" ]; then python -m pytest tests -x -q -m gpu 2>&1 | tail -3; fi
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
if [ -z "$QUICK" ]; then python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; echo "ref rc=$?"; fi
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
# the search kernels of the first two rounds of one frame (4 chains each), after the 3 warm-up frames
ncu --set full --clock-control none --import-source on -k k_trace --launch-skip 72 --launch-count 8 \
    -o /tmp/prof_${tag}_trace -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
echo "full k_trace rc=$?"
if [ -z "$QUICK" ]; then
ncu --set full --clock-control none -k regex:'k_trace8|k_finish|k_gen|k_shade_slots' --launch-skip 150 --launch-count 30 \
    -o /tmp/prof_${tag}_rest -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full_$tag.log 2>&1
echo "full rest rc=$?"
fi
for k in trace rest; do
  [ -f /tmp/prof_${tag}_$k.ncu-rep ] && ncu -i /tmp/prof_${tag}_$k.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_${k}_raw.csv 2>/dev/null
done
ls -la /tmp/prof_${tag}_*.ncu-rep
# keep one report with source for the hot-line view if it is small enough
true
du -sh gpurun_out
head -c 400 gpurun_out/bench_$tag.json; echo
