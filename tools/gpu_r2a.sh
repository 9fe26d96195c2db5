#!/bin/bash
# Round 2, call A: smoke, the GPU parity suite, A/B of the checkpoint store path of the fused march
# (DGADJ_TMA_STORE=0: STG per value, =1: bulk-TMA from the double-buffered park), default bench.
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== A/B store path"
for v in 0 1 0 1; do
  DGADJ_TMA_STORE=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ab_tma$v.json 2> $OUT/ab_tma$v.err
  echo "tma_store=$v rc=$? $(python -c "import json; d=json.load(open('$OUT/ab_tma$v.json')); r=d['roofline']; print('%.4e upd/s kern_ms %.2f frac %.3f alg %.3f smem %d' % (d['value'], r['kernel_ms'], r['frac'] or 0, r['frac_algorithmic'], d['plan']['smem_bytes']))" 2>&1 | tail -1)"
done
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "== bench (default)"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-1500 $OUT/bench.json; tail -3 $OUT/bench.err
