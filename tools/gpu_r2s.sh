#!/bin/bash
TAG=${1:-r2s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== tdg tests"; timeout 900 python -m pytest tests -q -m gpu -k "tdg or adaptive or reference_argument or cfg5" > $OUT/pytest_tdg.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|^E  " $OUT/pytest_tdg.log | cut -c1-250 | head -20
echo "== tdg / fd small batch"; timeout 600 python tools/bench_secondary.py tdg_fd > $OUT/tdg_fd.jsonl 2> $OUT/tdg_fd.err; echo "rc=$?"; cut -c1-420 $OUT/tdg_fd.jsonl; tail -3 $OUT/tdg_fd.err
