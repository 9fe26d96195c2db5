#!/bin/bash
# A/B of Burgers kernel builds (launch bounds variants).  usage: tools/gpu_burgers_ab.sh [tag]
TAG=${1:-r1bgab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
PKG=$PWD/adjoint-ode-adaptivity_b200
for v in "" _minb2 _minb3 _minb4; do
  DGADJ_LIB=$PKG/libdgadj$v.so timeout 600 python tools/bench_secondary.py burgers > $OUT/burgers$v.jsonl 2> $OUT/burgers$v.err
  echo "variant '$v' rc=$? $(python -c "
import json
print(' '.join('%.3e (%.1f ms)' % (json.loads(l)['value'], json.loads(l)['ms']) for l in open('$OUT/burgers$v.jsonl')))")"
done
