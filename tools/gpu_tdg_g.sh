#!/bin/bash
TAG=${1:-tdgg}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in 0 1; do
  DGADJ_TDG_VAR=$V timeout 300 python -m pytest tests -q -m gpu -k "tdg_warp or tdg_march" > $OUT/pytest_$V.log 2>&1; echo "rc=$?"; tail -1 $OUT/pytest_$V.log
  DGADJ_TDG_VAR=$V timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/l_$V.csv python tools/bench_secondary.py tdg_fd > $OUT/ncu_$V.log 2>&1
  echo "VAR=$V"; grep "tdg_march" $OUT/l_$V.csv | tail -3 | awk -F'","' '{print $5, $NF}' | cut -c1-160
done
