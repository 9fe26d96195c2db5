#!/bin/bash
TAG=${1:-bgf}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== burgers fused tests"; timeout 1200 python -m pytest tests -q -m gpu -k "burgers_fused or cfg3_full" > $OUT/pytest_bgf.log 2>&1; echo "rc=$?"; grep -E "passed|failed|AssertionError|^E  " $OUT/pytest_bgf.log | cut -c1-250 | head -30
echo "== timing"; timeout 900 python tools/bench_burgers_fused.py 16384 0.4 0,4 1,0 > $OUT/bgf_timing.jsonl 2> $OUT/bgf_timing.err; echo "rc=$?"; cut -c1-330 $OUT/bgf_timing.jsonl; tail -3 $OUT/bgf_timing.err
echo "== tdg / fd small batch"; timeout 600 python tools/bench_secondary.py tdg_fd > $OUT/tdg_fd.jsonl 2> $OUT/tdg_fd.err; echo "rc=$?"; cut -c1-400 $OUT/tdg_fd.jsonl; tail -3 $OUT/tdg_fd.err
echo "== ncu"
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:burgers_fused -s 1 -c 1 -o $OUT/prof_bgf python tools/bench_burgers_fused.py 1332 0.034 0 1 > $OUT/ncu_bgf.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_bgf.log
