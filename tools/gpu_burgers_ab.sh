#!/bin/bash
# A/B of Burgers kernel builds.  Variants are extra libraries built next to the default one, e.g.
#   make -C adjoint-ode-adaptivity_b200/csrc -j8 EXTRA="-D'BG_MINB_FWD(NP)=3'" OUT=../libdgadj_minb3.so BUILD=build_minb3
# usage: tools/gpu_burgers_ab.sh [tag] [variant suffixes...]
TAG=${1:-r1bgab}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
PKG=$PWD/adjoint-ode-adaptivity_b200
timeout 600 python -m pytest tests -q -m gpu -k "burgers or cfg3" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for v in "" "$@"; do
  DGADJ_LIB=$PKG/libdgadj$v.so timeout 600 python tools/bench_secondary.py burgers > $OUT/burgers$v.jsonl 2> $OUT/burgers$v.err
  echo "variant '$v' rc=$? $(python -c "
import json
print(' '.join('%.3e (%.1f ms)' % (json.loads(l)['value'], json.loads(l)['ms']) for l in open('$OUT/burgers$v.jsonl')))")"
done
