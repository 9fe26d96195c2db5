#!/usr/bin/env python
"""Instruction mix of an address range of one kernel in a cuobjdump -sass dump.
usage: sass_mix.py <sass file> <kernel substring> [<lo hex> <hi hex>] ; without a range prints
the loop ranges (backward branches) found in the kernel."""
import re, sys, collections
path, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else None
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else None
ins, on = [], False
for line in open(path):
    if 'Function :' in line:
        on = pat in line
        continue
    if not on:
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
if lo is None:
    for a, t in ins:
        m = re.search(r'BRA(?:\.U)?(?:\.ANY)?\s+(?:!?U?P\d,\s+)?(?:!?U?P\d,\s+)?0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            print(f"loop {int(m.group(1),16):#x} .. {a:#x}  ({(a-int(m.group(1),16))//16+1} instrs)")
    sys.exit(0)
c = collections.Counter()
for a, t in ins:
    if lo <= a <= hi:
        t = re.sub(r'^@!?U?P\d\s+', '', t)
        c[t.split()[0].split('.')[0]] += 1
tot = sum(c.values())
for k, v in c.most_common():
    print(f"{v:5d} {k}")
print(f"{tot:5d} total; fp64 pipe = {c['DFMA']+c['DMUL']+c['DADD']+c['DSETP']}")
