#!/bin/bash
# multi-GPU evidence: the NCCL C-ABI check, the nccl tests and the bench at N ranks (run under `gpurun --gpus N`)
N=${1:-2}; TAG=${2:-r2multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
echo "== nccl_abi_check x$N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tools/nccl_abi_check.py > $OUT/nccl_abi_check_N$N.log 2>&1; echo "rc=$?"; grep -E " ok|Error|error" $OUT/nccl_abi_check_N$N.log | head
echo "== bench x$N"
NCCL_DEBUG=INFO timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_N$N.json 2> $OUT/bench_N$N.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_N$N.json"))
print("N=%d value %.4e e2e %.4e frac %.3f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
c=d["secondary"]["cfg4_sweep_global"]; print("cfg4: %.4e upd/s ms %.1f sha %s" % (c["value"], c["ms"], c["norms_sha256"]))
PY
grep -v "NCCL INFO" $OUT/bench_N$N.err | tail -3 | cut -c1-300
grep -E "NCCL INFO.*Init COMPLETE" $OUT/bench_N$N.err | head -8 | cut -c1-200
echo "== pytest nccl"; timeout 600 python -m pytest tests -q -m gpu -k nccl > $OUT/pytest_nccl.log 2>&1; tail -3 $OUT/pytest_nccl.log
