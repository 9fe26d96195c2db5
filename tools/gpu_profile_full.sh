#!/bin/bash
# Full-size ncu capture of the dominant kernel on the bench configuration + launch list.
TAG=${1:-r1full}; OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > $OUT/plain.json 2> $OUT/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/plain2.json 2> $OUT/plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 1 -c 1 -o $OUT/prof_march_full $CMD > $OUT/ncu_full.log 2>&1
echo "full rc=$?"
ls -la $OUT; tail -3 $OUT/ncu_full.log
