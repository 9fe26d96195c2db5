#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into a JSON for profiles/: headline raw metrics, stall
reasons in total and per opcode.  usage: ncu_summary.py <rep> <out.json> "<command note>" """
import collections, csv, json, re, subprocess, sys
rep, out_path, note = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, d = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']
out = {}
for i, n in enumerate(h):
    if n in keep or n.startswith('smsp__average_warps_issue_stalled'):
        out[n] = {"unit": u[i], "value": d[i]}
mets = ",".join("smsp__pcsamp_warps_issue_stalled_" + m for m in
                ["math_pipe_throttle", "wait", "short_scoreboard", "mio_throttle", "barrier", "not_selected", "selected",
                 "long_scoreboard", "dispatch_stall", "lg_throttle"]) + ",inst_executed"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--metrics", mets],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
cols = hdr[2:]
tot = collections.Counter()
byop = collections.defaultdict(collections.Counter)
for r in data:
    t = re.sub(r'^@!?U?P\d\s+', '', r[1].strip())
    op = t.split()[0].split('.')[0]
    for c, v in zip(cols, r[2:]):
        try:
            byop[op][c] += int(float(v)); tot[c] += int(float(v))
        except ValueError:
            pass
T = sum(v for c, v in tot.items() if c.startswith('stall'))
out["_stall_samples_pct"] = {c: round(100 * v / T, 2) for c, v in tot.items() if c.startswith('stall')}
out["_by_opcode"] = {op: dict({c: round(100 * cnt[c] / T, 2) for c in cols if c.startswith('stall') and cnt[c]},
                              warp_insts_executed=cnt['Instructions Executed'])
                     for op, cnt in sorted(byop.items(), key=lambda kv: -kv[1]['Instructions Executed'])[:12]}
out["_command"] = note
json.dump(out, open(out_path, "w"), indent=1)
for k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread'):
    print(k, out.get(k))
print(out["_stall_samples_pct"])
