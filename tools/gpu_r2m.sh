#!/bin/bash
TAG=${1:-r2m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -k "adaptive_loop_advection or fused_randomised or edge_cases or handles_release" > $OUT/pytest.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|^E  " $OUT/pytest.log | cut -c1-300 | head -40
