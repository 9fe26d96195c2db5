#!/bin/bash
TAG=${1:-hp}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -k "hp or windowed or edge_cases or abi or adaptive_loop_advection" > $OUT/pytest.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|^E  " $OUT/pytest.log | cut -c1-250 | head -20
