#!/bin/bash
# instrumented library variant for tools/wave_latency.py (per-ray stage timestamps in k_wave); never the product build
mkdir -p build_variants
CGRT_LIB=$PWD/build_variants/lib_lat.so CGRT_NVCC_EXTRA="-DCGRT_WAVE_LAT ${1}" python -c "import __graft_entry__ as g; g.build(force=True)"
