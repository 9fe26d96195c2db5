#!/bin/bash
TAG=${1:-r2i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== adaptive loop + tdg tests"; timeout 1200 python -m pytest tests -q -m gpu -k "adaptive or tdg or fd_path or randomised_secondary" > $OUT/pytest_adapt.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|^E  " $OUT/pytest_adapt.log | cut -c1-250 | head -40
echo "== cfg5 timing"; timeout 600 python - > $OUT/cfg5.json 2> $OUT/cfg5.err <<PY
import sys, json, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import torch, dgadj_loader, secondary
pkg = dgadj_loader.load_package()
print(json.dumps(secondary.cfg5(pkg, torch, torch.device("cuda", 0)), indent=1))
PY
echo "rc=$?"; cat $OUT/cfg5.json | head -60; tail -5 $OUT/cfg5.err
