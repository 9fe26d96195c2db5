#!/bin/bash
# the GPU suite and the default bench line (a shorter form of gpu_final.sh)
TAG=${1:-suite}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log | cut -c1-200
echo "== bench (default)"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
r=d["roofline"]
print("value %.4e e2e %.4e frac %.3f" % (d["value"], d["e2e"]["value"], r["frac"]))
for k,v in d["secondary"].items():
    if k=="cfg5_adaptive": print(k, v["tdg"]["ms_per_iteration"], v["fd"]["ms_per_iteration"])
    else: print(k, "%.4e" % v["value"], v["roofline"]["frac"])
PY
