#!/bin/bash
TAG=${1:-r2h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench (default)"; timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.4e e2e %.4e frac %.3f alg %.3f launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["frac_algorithmic"], d["gpu_launches"]))
print(json.dumps(d.get("secondary"), indent=1)[:6000])
PY
tail -5 $OUT/bench.err
echo "== pytest -m gpu (all)"; timeout 1800 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
