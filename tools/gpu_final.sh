#!/bin/bash
# Round-end evidence on one GPU: smoke, the whole GPU suite, the default bench line, the ncu launch list of the
# same bench command (gpu__time_duration per launch), the fused Burgers timing.
TAG=${1:-r2final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
echo "== bench (default)"; timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
r=d["roofline"]
print("value %.4e e2e %.4e frac %.3f alg %.3f launches %d kernel_ms %.1f traffic %s" % (d["value"], d["e2e"]["value"], r["frac"], r["frac_algorithmic"], d["gpu_launches"], r["kernel_ms"], r["traffic_source"]))
for k,v in d["secondary"].items():
    if k=="cfg5_adaptive": print(k, v["tdg"]["ms_per_iteration"], v["fd"]["ms_per_iteration"], v["fd"].get("cpu_baseline",{}).get("value"))
    else: print(k, "%.4e" % v["value"], v["roofline"]["frac"], v.get("cpu_baseline",{}).get("value"), v.get("max_rel_diff_vs_oracle"))
PY
tail -3 $OUT/bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-200 $OUT/bench_ref.json
echo "== burgers timing"; timeout 600 python tools/bench_burgers_fused.py 16384 0.4 0 1,0 > $OUT/bgf_timing.jsonl 2> $OUT/bgf_timing.err; cut -c1-260 $OUT/bgf_timing.jsonl
echo "== ncu launch list"
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary"
$CMD > $OUT/plain_list.json 2> $OUT/plain_list.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -4 $OUT/launches.csv | cut -c1-220
