#!/bin/bash
# Burgers parity tests + secondary bench of config 3 (one gpurun call).  usage: tools/gpu_burgers.sh [tag]
TAG=${1:-r1bg}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -q -m gpu -k "burgers or limiter or Burgers" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
timeout 600 python tools/bench_secondary.py burgers > $OUT/burgers.jsonl 2> $OUT/burgers.err; echo "bench rc=$?"; cut -c1-80,200-420 $OUT/burgers.jsonl
