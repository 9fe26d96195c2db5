#!/bin/bash
TAG=${1:-small_batch}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== fd / adaptive tests"; timeout 900 python -m pytest tests -q -m gpu -k "fd or tdg or adaptive or reference_argument or cfg5 or caches or matlab or time_dg" > $OUT/pytest_fd.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|^E  " $OUT/pytest_fd.log | cut -c1-250 | head -20
echo "== tdg / fd small batch"; timeout 600 python tools/bench_secondary.py tdg_fd > $OUT/tdg_fd.jsonl 2> $OUT/tdg_fd.err; echo "rc=$?"; cut -c1-300 $OUT/tdg_fd.jsonl; tail -3 $OUT/tdg_fd.err
timeout 600 python - > $OUT/cfg5.json 2> $OUT/cfg5.err <<PY
import sys, json, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import torch, dgadj_loader, secondary
pkg = dgadj_loader.load_package()
r = secondary.cfg5(pkg, torch, torch.device("cuda", 0))
print(json.dumps({k: {kk: vv for kk, vv in v.items() if "ms_per" in kk} for k, v in r.items() if isinstance(v, dict)}))
PY
cat $OUT/cfg5.json; tail -3 $OUT/cfg5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_tdg_fd.csv python tools/bench_secondary.py tdg_fd > $OUT/ncu.log 2>&1; echo "rc=$?"
python - <<PY
import csv,collections,re
rows=[r for r in csv.reader(open("$OUT/launches_tdg_fd.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: v=float(r[ix['Metric Value']].replace(',',''))
    except: continue
    name=re.sub(r'\(.*','',r[ix['Kernel Name']])[:70]+" grid "+r[ix['Grid Size']]+" block "+r[ix['Block Size']]
    agg[name].append(v*1e-3)
for k,v in agg.items():
    if 'dgadj' in k: print("%-110s x%-3d min %.1f us median %.1f us"%(k,len(v),min(v),sorted(v)[len(v)//2]))
PY
