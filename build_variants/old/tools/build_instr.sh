#!/bin/bash
# instrumented A/B build of the library (scheduling counters + timeline), kept out of the product path
mkdir -p build_variants
CGRT_LIB=$PWD/build_variants/lib_instr.so CGRT_NVCC_EXTRA="-DCGRT_INSTRUMENT ${1}" python -c "import __graft_entry__ as g; g.build(force=True)"
