#!/bin/bash
TAG=${1:-r2t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
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
