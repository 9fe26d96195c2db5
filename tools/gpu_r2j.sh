#!/bin/bash
TAG=${1:-r2j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -k "reference_argument or matlab_named" > $OUT/pytest.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|^E  " $OUT/pytest.log | cut -c1-300 | head -40
