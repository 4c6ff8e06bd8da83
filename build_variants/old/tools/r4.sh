CGRT_LIB=$PWD/build_variants/lib_nopf.so ncu --set full --clock-control none --import-source on -k 'regex:k_trace' --launch-skip 24 --launch-count 6 -o gpurun_out/prof_trace -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_trace.log 2>&1
echo rc=$?
