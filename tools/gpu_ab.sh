#!/bin/bash
# A/B of library variants on config 3 (fused Burgers): tools/gpu_ab.sh tag lib1 lib2 ...
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for lib in "$@"; do
  L=$PWD/adjoint-ode-adaptivity_b200/$lib
  [ -f $L ] || { echo "$lib missing"; continue; }
  DGADJ_LIB=$L timeout 600 python tools/bench_burgers_fused.py 16384 0.4 0 1,0 > $OUT/ab_$lib.jsonl 2> $OUT/ab_$lib.err
  echo "$lib rc=$? $(python - <<PY
import json
for l in open("$OUT/ab_$lib.jsonl"):
    d=json.loads(l); print("ind=%d %.4e upd/s (%.0f ms)" % (d["indicator"], d["value"], d["ms"]), end="; ")
PY
)"
done
