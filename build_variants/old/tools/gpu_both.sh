#!/bin/bash
# GPU regression in both traversal modes, then the short bench summary
python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "replayed|passed|failed|Error|error|assert" | tail -15
CGRT_EXACT_ONLY=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
bash tools/gpu_check.sh ${1:-10} | tail -2
