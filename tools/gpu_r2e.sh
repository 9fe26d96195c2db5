#!/bin/bash
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== burgers fused tests"; timeout 1200 python -m pytest tests -q -m gpu -k "burgers_fused" > $OUT/pytest_bgf.log 2>&1; echo "rc=$?"; grep -E "passed|failed|AssertionError|^E  " $OUT/pytest_bgf.log | cut -c1-250 | head -60
