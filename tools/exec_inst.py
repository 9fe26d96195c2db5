#!/usr/bin/env python
"""fp64-pipe instructions the fused march kernel EXECUTES per element-stage update, counted from the
SASS of the built library (no GPU needed):

    python tools/exec_inst.py [--np 9] [--ept 4] [--bd 256] [--out profiles/exec_inst.json]

The kernel's time loops are `step loop { stage loop x nstages }` (forward phase: a fine and a coarse
stage loop per step; adjoint phase: one stage loop per step).  Loops are recognised by their backward
branches; a stage loop is a loop with >= 20 fp64 instructions and no other such loop inside, a step loop
its smallest enclosing loop.  Per thread and step the count is  nstages x body(stage loops) (the rest of
a step-loop body is left out: see analyse());  per update it is that over EPT elements x nstages stages
x 2 (one forward and one adjoint update per element-stage).  fp64-pipe opcodes: DFMA DMUL DADD DSETP (and DMNMX / F2F.F64 if any).
bench.py reads the JSON (stamped with the sha256 of the kernel source) for `roofline.frac`.
"""
import argparse
import collections
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "adjoint-ode-adaptivity_b200", "csrc")
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def kernel_source_sha():
    h = hashlib.sha256()
    for f in ("dgadj_kernels.cuh", "dgadj_march_np.cu"):
        h.update(open(os.path.join(CSRC, f), "rb").read())
    return h.hexdigest()[:16]


def parse_sass(text):
    ins = []
    for line in text.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            t = re.sub(r"^@!?U?P\d\s+", "", m.group(2).strip())
            ins.append((int(m.group(1), 16), t))
    return ins


def loops_of(ins):
    out = []
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?(?:\.ANY)?\s+(?:!?U?P\d,\s+)?(?:!?U?P\d,\s+)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            out.append((int(m.group(1), 16), a))
    return out


def count(ins, lo, hi, ops=FP64):
    c = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi:
            c[t.split()[0].split(".")[0]] += 1
    return sum(c[o] for o in ops), sum(c.values()), c


def analyse(ins, ept, nstages=5):
    loops = loops_of(ins)
    # candidates: loops that hold real arithmetic (>= 20 fp64 instructions) -- not the spin loops of the waits
    big = [(lo, hi) for lo, hi in loops if count(ins, lo, hi)[0] >= 20]
    # spin loops of mbarrier waits branch back from an out-of-line tail: they "contain" later code; drop loops
    # that contain the start of another big loop's *enclosing* step loop by keeping only properly nested ones
    stage = [l for l in big if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in big)]
    res = []
    for lo, hi in stage:
        parents = [o for o in big if o != (lo, hi) and o[0] <= lo and hi <= o[1]]
        parent = min(parents, key=lambda o: o[1] - o[0]) if parents else None
        res.append(dict(stage=(lo, hi), step=parent))
    steps = collections.OrderedDict()
    for r in res:
        steps.setdefault(r["step"], []).append(r["stage"])
    report, total_fp64, total_all = [], 0, 0
    for step, stages in steps.items():
        if step is None:
            continue
        f_step, n_step, _ = count(ins, *step)
        f_st = [count(ins, *s)[0] for s in stages]
        n_st = [count(ins, *s)[1] for s in stages]
        # stage loops only: the rest of a step-loop body is mostly conditional code the bench does not
        # run (the history write with its dense V products), so a static count of it would overstate
        # the executed work; what is left out (forming rho, the indicator dot product: 2 fp64
        # instructions per update) makes the figure a slight UNDER-count -- conservative for `frac`
        fp = nstages * sum(f_st)
        al = nstages * sum(n_st)
        total_fp64 += fp
        total_all += al
        report.append(dict(step_loop=[hex(step[0]), hex(step[1])], stage_loops=[[hex(a), hex(b)] for a, b in stages],
                           fp64_per_stage_loop_pass=f_st, issued_per_stage_loop_pass=n_st,
                           fp64_per_thread_step=fp, issued_per_thread_step=al))
    per_update = total_fp64 / (ept * nstages * 2.0)
    return dict(loops=report, fp64_inst_per_update=per_update, issued_inst_per_update=total_all / (ept * nstages * 2.0),
                fp64_share_of_issued=total_fp64 / max(total_all, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=9)
    ap.add_argument("--ept", type=int, default=4)
    ap.add_argument("--bd", type=int, default=256)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "exec_inst.json"))
    ap.add_argument("--sass-out", default=None, help="also write the kernel's SASS listing here")
    a = ap.parse_args()
    obj = os.path.join(CSRC, "build", f"dgadj_march_np{a.np}.o")
    fun = f"_ZN5dgadj12march_kernelILi{a.np}ELi{a.ept}ELi{a.bd}ELb1ELb1ELb1ELb0EEEvNS_5KArgsE"
    text = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], stdout=subprocess.PIPE, text=True, check=True).stdout
    if a.sass_out:
        open(a.sass_out, "w").write(text)
    ins = parse_sass(text)
    if not ins:
        sys.exit(f"no SASS for {fun} in {obj}")
    r = analyse(ins, a.ept)
    _, _, hist = count(ins, 0, 1 << 30)
    key = f"np{a.np}_ept{a.ept}_bd{a.bd}_fused"
    try:
        doc = json.load(open(a.out))
    except Exception:
        doc = {}
    doc[key] = dict(kernel=fun, kernel_source_sha=kernel_source_sha(), **r,
                    opcode_histogram_static={k: v for k, v in hist.most_common(40)},
                    blackwell_markers={k: hist.get(k, 0) for k in ("UBLKCP", "SYNCS", "DFMA", "LDCU", "UTMASTG", "STG", "LDG", "STS", "LDS")})
    doc["_how"] = "tools/exec_inst.py: static SASS count weighted by the loop trip counts (5 stages per step); see its docstring"
    json.dump(doc, open(a.out, "w"), indent=1)
    print(f"{key}: fp64 {r['fp64_inst_per_update']:.1f} / issued {r['issued_inst_per_update']:.1f} per update "
          f"(fp64 share {r['fp64_share_of_issued']:.3f}); sha {kernel_source_sha()}")


if __name__ == "__main__":
    main()
