#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: total time, share, launches.
usage: launch_list_summary.py <launches.csv> <out.txt> "<command / call note>" """
import collections
import csv
import re
import sys

src, out_path, note = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ix["Metric Unit"]], 1e-6)
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:90]
    agg[name][0] += ms
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
n = sum(v[1] for v in agg.values())
out = [note, "launches: %d, total %.1f ms (cold-cache, serialised by the profiler: shares matter, not absolutes)" % (n, tot)]
for k, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.append("%8.2f ms  %5.1f %%  x%-3d %s" % (ms, 100 * ms / tot, c, k))
march = sum(v[0] for k, v in agg.items() if "march_kernel" in k)
peak = sum(v[0] for k, v in agg.items() if "dfma_peak" in k)
if march:
    out.append("march_kernel share of the whole command: %.1f %%; of the step kernels (without the DFMA peak microbenchmark that "
               "bench.py runs after the timed region): %.1f %%" % (100 * march / tot, 100 * march / (tot - peak)))
open(out_path, "w").write("\n".join(out) + "\n")
print("\n".join(out[:5] + out[-1:]))
