#!/bin/bash
# Round 2, call B: A/B of the two builds of the checkpoint store path (libdgadj.so: STG; libdgadj_tma.so:
# -DDGADJ_TMA_STORE_PATH=1, bulk-TMA from a double-buffered park), GPU suite, ncu --set full of both.
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
PKG=adjoint-ode-adaptivity_b200
for rep in 1 2; do
for v in stg tma; do
  LIB=$PWD/$PKG/libdgadj.so; [ $v = tma ] && LIB=$PWD/$PKG/libdgadj_tma.so
  DGADJ_LIB=$LIB timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ab_${v}_$rep.json 2> $OUT/ab_${v}_$rep.err
  echo "$v rc=$? $(python -c "import json; d=json.load(open('$OUT/ab_${v}_$rep.json')); r=d['roofline']; print('%.4e upd/s kern_ms %.2f alg %.3f smem %d' % (d['value'], r['kernel_ms'], r['frac_algorithmic'], d['plan']['smem_bytes']))" 2>&1 | tail -1)"
done; done
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "== pytest fused tests on the TMA build"; DGADJ_LIB=$PWD/$PKG/libdgadj_tma.so timeout 900 python -m pytest tests -q -m gpu -x -k "fused or cfg2 or windowed or edge" > $OUT/pytest_gpu_tma.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_tma.log
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
for v in stg tma; do
  LIB=$PWD/$PKG/libdgadj.so; [ $v = tma ] && LIB=$PWD/$PKG/libdgadj_tma.so
  DGADJ_LIB=$LIB timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 1 -c 1 -o $OUT/prof_march_$v $CMD > $OUT/ncu_$v.log 2>&1
  echo "ncu $v rc=$?"; tail -2 $OUT/ncu_$v.log
done
ls -la $OUT
