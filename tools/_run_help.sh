run() { CGRT_WAVE="fin=$2,settle=$3" python tools/wave_prof.py dragon 1920 1080 5 8 0 $1 2>&1 | tail -2 | head -1 | sed 's/.*device ms//' | cut -c1-100; }
for f in 6 7 8 10; do echo -n "settle fin=$f: "; run 1 $f 1; done
for f in 6 7; do echo -n "no settle fin=$f: "; run 1 $f 0; done
