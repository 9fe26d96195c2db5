#!/usr/bin/env python
"""Config 3 through the fused Burgers kernel (dgadj_burgers_fwd_adj): B=16384, N=4, K=256, T past shock
formation.  One JSON line per variant.  usage: bench_burgers_fused.py [B] [T] [ept list] [ind list]"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import dgadj_loader

pkg = dgadj_loader.load_package()
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
epts = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
inds = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1]
N, K = 4, 256
s = pkg.BurgersDG1D(N, K, domain=(-1.0, 1.0), bc="periodic")
g = torch.Generator(device=dev); g.manual_seed(1235)
c = torch.rand((B, 1, 1), dtype=torch.float64, device=dev, generator=g) - 0.5
A = torch.rand((B, 1, 1), dtype=torch.float64, device=dev, generator=g) + 0.5
ph = torch.rand((B, 1, 1), dtype=torch.float64, device=dev, generator=g) * 2 * math.pi
x = torch.tensor(s.g.x, device=dev)[None]
u0 = (c + A * torch.sin(math.pi * x + ph)).contiguous()
dt = s.stable_dt(2.0)
S = int(math.ceil(T / dt))
reps = int(os.environ.get("REPS", 2))
for ind in inds:
    for ept in epts:
        s.set_tuning(elems_per_thread=ept)
        out = s.fwd_adj(u0, dt, S, indicator=bool(ind), want_uT=False, want_lam0=False)     # warm-up (allocates the ring)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = s.fwd_adj(u0, dt, S, indicator=bool(ind), want_uT=False, want_lam0=False)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        ups = 2 * 5 * S * K * B
        print(json.dumps(dict(workload="config 3 fused: Burgers + SlopeLimitN N=4 K=256 B=%d S=%d (T=%.3f), fwd + adjoint%s" %
                              (B, S, S * dt, " (enriched) + indicator" if ind else " (coarse)"),
                              indicator=ind, value=ups / (best * 1e-3), unit="updates/s", ms=best, plan=s.plan(B, bool(ind)),
                              limited_fraction=float(out["nlim"][:, 0].double().mean()) / (5 * S * K),
                              status_max=int(out["status"].max()))), flush=True)
