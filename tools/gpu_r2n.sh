#!/bin/bash
TAG=${1:-r2n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -k "per_trajectory" > $OUT/pytest.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|^E  " $OUT/pytest.log | cut -c1-300 | head -40
