#!/bin/bash
# Full-size ncu evidence of the dominant kernel on the bench configuration, one ncu run per call:
#   tools/gpu_profile_full.sh list [tag]   launch list (gpu__time_duration per launch)
#   tools/gpu_profile_full.sh full [tag]   --set full capture of one march_kernel launch
MODE=${1:-list}; TAG=${2:-r1full}; OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > $OUT/plain_$MODE.json 2> $OUT/plain_$MODE.err || { echo "plain run failed"; tail -5 $OUT/plain_$MODE.err; exit 1; }
if [ "$MODE" = list ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
  echo "launch list rc=$?"; tail -3 $OUT/launches.csv | cut -c1-200
else
  ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 1 -c 1 -o $OUT/prof_march_full $CMD > $OUT/ncu_full.log 2>&1
  echo "full rc=$?"; tail -3 $OUT/ncu_full.log
fi
ls -la $OUT
