#!/bin/bash
TAG=${1:-extra}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== tdg / fd large batch"; SEC_B=262144 timeout 600 python tools/bench_secondary.py tdg_fd > $OUT/tdg_fd_large.jsonl 2> $OUT/tdg_fd_large.err; echo "rc=$?"; cut -c1-330 $OUT/tdg_fd_large.jsonl; tail -2 $OUT/tdg_fd_large.err
echo "== ncu burgers post-shock"
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:burgers_fused -s 1 -c 1 -o $OUT/prof_bgf_postshock python tools/bench_burgers_fused.py 1332 0.4 0 1 > $OUT/ncu_bgf.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_bgf.log
