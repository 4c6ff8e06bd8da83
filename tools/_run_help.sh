run() { CGRT_WAVE=fin=$2 python tools/wave_prof.py dragon 1920 1080 5 8 0 $1 2>&1 | tail -2 | head -1 | sed 's/.*device ms//' | cut -c1-100; }
echo -n "product 1/1: "; run 1 6
for v in WAVE_STEPS_4 WAVE_STEPS_16 WAVE_STEPS_32 FAST_W_LEAF_2; do echo -n "$v 1/1: "; CGRT_LIB=$PWD/build_variants/lib_$v.so run 1 6; done
echo -n "product 1/1: "; run 1 6
