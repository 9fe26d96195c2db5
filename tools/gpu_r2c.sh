#!/bin/bash
# Round 2, call C: first GPU run of the fused Burgers kernel (tests, timing at full size, ncu at reduced size)
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== burgers fused tests"; timeout 1200 python -m pytest tests -q -m gpu -x -k "burgers_fused or cfg3_full" > $OUT/pytest_bgf.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_bgf.log
echo "== timing (default build: 4 CTAs/SM at 128 registers)"; timeout 900 python tools/bench_burgers_fused.py 16384 0.4 0 1,0 > $OUT/bgf_timing.jsonl 2> $OUT/bgf_timing.err; echo "rc=$?"; cut -c1-400 $OUT/bgf_timing.jsonl; tail -3 $OUT/bgf_timing.err
if [ -f adjoint-ode-adaptivity_b200/libdgadj_b3.so ]; then
echo "== timing (3 CTAs/SM at 168 registers)"; DGADJ_LIB=$PWD/adjoint-ode-adaptivity_b200/libdgadj_b3.so timeout 900 python tools/bench_burgers_fused.py 16384 0.4 0 1,0 > $OUT/bgf_timing_b3.jsonl 2> $OUT/bgf_timing_b3.err; echo "rc=$?"; cut -c1-400 $OUT/bgf_timing_b3.jsonl; tail -3 $OUT/bgf_timing_b3.err
fi
echo "== ncu"
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:burgers_fused -s 1 -c 1 -o $OUT/prof_bgf python tools/bench_burgers_fused.py 1332 0.034 0 1 > $OUT/ncu_bgf.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_bgf.log
ls -la $OUT
