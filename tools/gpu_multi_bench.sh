#!/bin/bash
N=${1:-2}; TAG=${2:-r2multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench x$N"
NCCL_DEBUG=INFO timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_N$N.json 2> $OUT/bench_N$N.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_N$N.json"))
print("N=%d value %.4e e2e %.4e frac %.3f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
c=d["secondary"]["cfg4_sweep_global"]; print("cfg4: %.4e upd/s ms %.1f sha %s" % (c["value"], c["ms"], c["norms_sha256"]))
print({k:(v.get("value") if isinstance(v,dict) else v) for k,v in d["secondary"].items()})
PY
grep -v "NCCL INFO" $OUT/bench_N$N.err | tail -5 | cut -c1-300
grep -E "NCCL INFO.*(Init COMPLETE|NVLS multicast)" $OUT/bench_N$N.err | head -4 | cut -c1-250
