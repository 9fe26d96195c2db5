#!/bin/bash
# Fused fwd+adjoint throughput over orders and mesh sizes (reduced batch, device-resident leg only).
# usage: tools/gpu_shape_sweep.sh [tag]
TAG=${1:-r1shape}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "N K B updates/s frac_of_fp64_peak(algorithmic) ept block grid" > $OUT/table.txt
for N in 1 2 3 4 6 8; do
  for K in 64 256 1000 1024; do
    B=$(( 148 * 1024 * 16 / K ))
    timeout 300 python bench.py --N $N --K $K --B $B --S 50 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/n${N}_k${K}.json 2> $OUT/n${N}_k${K}.err
    python - <<PY >> $OUT/table.txt
import json
try:
    d = json.load(open("$OUT/n${N}_k${K}.json")); r = d["roofline"]; p = d["plan"]
    print($N, $K, $B, "%.3e" % d["value"], "%.3f" % r["frac"], p["elems_per_thread"], p["block"], p["grid"])
except Exception as e:
    print($N, $K, $B, "FAILED", e)
PY
  done
done
cat $OUT/table.txt
