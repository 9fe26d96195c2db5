#!/bin/bash
# A/B of library builds / launch shapes on a reduced batch.  usage: tools/gpu_ab.sh tag
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { # name env... -- args
  name=$1; shift
  timeout 300 env "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/$name.json')); r=d['roofline']
    print('%-28s %.3e upd/s  frac %.3f  kern_ms %.2f  peak %.1f TF  clk %s' % ('$name', d['value'], r['frac'], r['kernel_ms'], r['peak'], d['clocks']['sm_mhz']))
except Exception as e:
    print('$name FAILED', e); print(open('$OUT/$name.err').read()[-600:])
PY
}
B="--B 4736 --S 50 --steps 2 --warmup 1 --no-e2e --no-cpu"
PKG=adjoint-ode-adaptivity_b200
timeout 600 python -m pytest tests -q -m gpu -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
run split_ept2_b512      DGADJ_X=1 python bench.py $B
run nosplit_ept2_b512    DGADJ_LIB=$PWD/$PKG/libdgadj_nosplit.so python bench.py $B
run split_ept1_b1024     DGADJ_X=1 python bench.py $B --ept 1 --block 1024
run nosplit_ept1_b1024   DGADJ_LIB=$PWD/$PKG/libdgadj_nosplit.so python bench.py $B --ept 1 --block 1024
run split_ept2_K512_b256 DGADJ_X=1 python bench.py --B 9472 --S 50 --steps 2 --warmup 1 --no-e2e --no-cpu --K 512
run split_ept2_K512_b512 DGADJ_X=1 python bench.py --B 9472 --S 50 --steps 2 --warmup 1 --no-e2e --no-cpu --K 512 --block 512
run split_N4_K64         DGADJ_X=1 python bench.py --B 151552 --S 50 --steps 2 --warmup 1 --no-e2e --no-cpu --N 4 --K 64
