#!/bin/bash
# Full evidence pass on one B200 (run under gpurun): parity tests, contract bench (both arms), ncu launch list, ncu --set full
# usage: tools/gpu_profile.sh <tag>   (outputs under gpurun_out/, kept below the 64 MiB pull limit: raw CSV pages, small reps)
tag=${1:-r02x}
mkdir -p gpurun_out
if [ -z "$QUICK" ]; then python -m pytest tests -x -q -m gpu 2>&1 | tail -3; fi
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
if [ -z "$QUICK" ]; then python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; echo "ref rc=$?"; fi
# launch list of the bench command: every launch with its device time (the persistent wavefront is ONE launch per frame, so
# its time here - cold cache, serialised - can be compared with the bench's ms_per_step directly)
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
# the frame kernel, two launches after the warm-up frames
ncu --set full --clock-control none --import-source on -k regex:k_wave --launch-skip 4 --launch-count 2 \
    -o gpurun_out/prof_${tag}_wave -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
echo "full k_wave rc=$?"
[ -f gpurun_out/prof_${tag}_wave.ncu-rep ] && ncu -i gpurun_out/prof_${tag}_wave.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_wave_raw.csv 2>/dev/null
du -sh gpurun_out
head -c 400 gpurun_out/bench_$tag.json; echo
