#!/usr/bin/env python
"""Executed warp instructions per CUDA source line: joins `nvdisasm -g` of one kernel (line info) with
the SASS page of an ncu report of the same build (same instruction order).
usage: sass_by_line.py <kernel.sass from nvdisasm -g> <ncu sass csv> <source.cu> <warp-updates> [top]"""
import collections, csv, re, sys
sass, ncsv, srcf, upd = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
ins, cur = [], 0
for l in open(sass):
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append((cur, m.group(2).strip()))
rows = list(csv.reader(open(ncsv)))[2:]
assert len(ins) == len(rows), (len(ins), len(rows))
src = open(srcf).read().split("\n")
byline, byop = collections.Counter(), collections.defaultdict(collections.Counter)
tot = 0
for (ln, t), r in zip(ins, rows):
    op = re.sub(r'^@!?U?P\d\s+', '', t).split()[0].split('.')[0]
    try:
        v = int(float(r[2]))
    except ValueError:
        v = 0
    byline[ln] += v; byop[ln][op] += v; tot += v
print("total per warp-update %.1f" % (tot / upd))
for ln, v in byline.most_common(top):
    print("%6.1f  %4d  %-88s %s" % (v / upd, ln, src[ln - 1].strip()[:88], dict((k, round(c / upd, 1)) for k, c in byop[ln].most_common(4))))
