#!/bin/bash
TAG=${1:-r2o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -k "hp_per_element or fused_fwd_adj_indicator or cfg2_size or edge_cases or burgers_fused_modes" > $OUT/pytest.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|^E  " $OUT/pytest.log | cut -c1-300 | head -40
timeout 600 python - > $OUT/cfg5.json 2> $OUT/cfg5.err <<PY
import sys, json, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import torch, dgadj_loader, secondary
pkg = dgadj_loader.load_package()
r = secondary.cfg5(pkg, torch, torch.device("cuda", 0))
print(json.dumps({k: {kk: vv for kk, vv in v.items() if "ms_per" in kk} for k, v in r.items() if isinstance(v, dict)}))
PY
cat $OUT/cfg5.json; tail -3 $OUT/cfg5.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > $OUT/bench_quick.json 2> $OUT/bench_quick.err; python -c "
import json; d=json.load(open('$OUT/bench_quick.json')); print('bench %.4e frac %.3f' % (d['value'], d['roofline']['frac']))"
