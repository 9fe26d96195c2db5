#!/bin/bash
TAG=${1:-cfg5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for rep in 1 2 3; do
timeout 600 python - > $OUT/cfg5_$rep.json 2> $OUT/cfg5_$rep.err <<PY
import sys, json, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import torch, dgadj_loader, secondary
pkg = dgadj_loader.load_package()
r = secondary.cfg5(pkg, torch, torch.device("cuda", 0))
print(json.dumps({k: {kk: vv for kk, vv in v.items() if "ms_per" in kk} for k, v in r.items() if isinstance(v, dict)}))
PY
cat $OUT/cfg5_$rep.json; tail -2 $OUT/cfg5_$rep.err
done
nproc; uptime
