#!/bin/bash
# One gpurun call: smoke, the GPU parity suite, a launch-shape sweep at reduced batch and the
# default bench.  Everything is logged under gpurun_out/.  usage: tools/gpu_check.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
echo "== sweep"
for cfg in "1 1024" "2 512" ; do
  set -- $cfg
  timeout 300 python bench.py --B 2368 --S 50 --steps 2 --warmup 1 --no-e2e --no-cpu --ept $1 --block $2 > $OUT/sweep_ept$1_b$2.json 2> $OUT/sweep_ept$1_b$2.err
  echo "ept=$1 block=$2 rc=$? $(python -c "import json,sys; d=json.load(open('$OUT/sweep_ept$1_b$2.json')); print('%.3e upd/s frac %.3f kern_ms %.1f peak %.1f TF' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['peak']))" 2>&1 | tail -1)"
done
echo "== bench (default)"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-300 $OUT/bench_ref.json
