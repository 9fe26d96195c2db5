#!/bin/bash
# All secondary workloads at their configured sizes + the time-DG / FD pair at a
# batch that fills the GPU.  usage: tools/gpu_secondary.sh [tag]
TAG=${1:-r1sec}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python tools/bench_secondary.py forward long_march sweep burgers tdg_fd > $OUT/secondary.jsonl 2> $OUT/secondary.err; echo "rc=$?"
SEC_B=262144 timeout 600 python tools/bench_secondary.py tdg_fd > $OUT/secondary_bigB.jsonl 2>> $OUT/secondary.err; echo "rc=$?"
python - <<PY
import json
for f in ("$OUT/secondary.jsonl", "$OUT/secondary_bigB.jsonl"):
    for l in open(f):
        d = json.loads(l); print("%.4e  %8.2f ms  %s" % (d["value"], d["ms"], d["workload"][:90]))
PY
