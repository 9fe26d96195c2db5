#!/bin/bash
TAG=${1:-tdgn}
OUT=gpurun_out/$TAG
mkdir -p $OUT
DGADJ_TDG_VAR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tdg_march_group -s 2 -c 1 -o $OUT/prof_tdg_march python tools/bench_secondary.py tdg_fd > $OUT/ncu.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu.log
