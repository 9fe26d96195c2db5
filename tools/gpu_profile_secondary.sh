#!/bin/bash
# ncu --set full capture of one secondary kernel at reduced size (one ncu run per gpurun call).
# usage: tools/gpu_profile_secondary.sh <workload: burgers|sweep|tdg_fd> <kernel regex> [tag] [skip]
W=${1:-burgers}; RE=${2:-burgers_kernel}; TAG=${3:-r1sec}; SKIP=${4:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export SEC_B=${SEC_B:-4096} SEC_S=${SEC_S:-50}
CMD="python tools/bench_secondary.py $W"
$CMD > $OUT/plain_$W.jsonl 2> $OUT/plain_$W.err || { echo "plain run failed"; tail -5 $OUT/plain_$W.err; exit 1; }
cat $OUT/plain_$W.jsonl | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c 1 -o $OUT/prof_$W $CMD > $OUT/ncu_$W.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu_$W.log; ls -la $OUT
