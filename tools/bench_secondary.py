#!/usr/bin/env python
"""Secondary workloads of SURVEY section 8(d) (not the driver's headline; numbers go to profiles/):
  forward the forward march alone (the reference's own workload), config 2 mesh
  long_march  S = 20000 steps on the config 2 mesh through dgadj_fwd_adj_windowed
  sweep   config 4: parameter sweep, N=4, K=64, S=100, per-trajectory speed a and CFL dt
  burgers config 3: Burgers + SlopeLimitN, N=4, K=256, B=16384, forward + checkpoints
  tdg     config 5: DG-in-time march + adjoint (u' = sin u), B=4096 ICs, refined mesh
  fd      config 5: finite-difference path, same ICs
One JSON line per workload.  CUDA events on the launching stream, 3 warm-ups, best of 5."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import dgadj_loader

pkg = dgadj_loader.load_package()
dev = torch.device("cuda", 0)
TWO_PI = 2 * math.pi


def timeit(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def sweep(B=131072):
    N, K, S = 4, 64, 100
    s = pkg.AdvecDG1D(N, K, domain=(0.0, TWO_PI), alpha=0.0, bc="periodic")
    n1 = int(round(math.sqrt(B)))
    a = torch.linspace(0.5, 2.0, n1, dtype=torch.float64, device=dev) * TWO_PI
    sig = torch.linspace(0.02, 0.2, B // n1, dtype=torch.float64, device=dev)
    A, SG = torch.meshgrid(a, sig, indexing="ij")
    A, SG = A.reshape(-1).contiguous(), SG.reshape(-1).contiguous()
    B = A.numel()
    x = torch.tensor(s.g.x, device=dev)[None]
    u0 = torch.exp(-(x - math.pi) ** 2 / (2 * SG[:, None, None] ** 2)).contiguous()
    dt0, _ = s.cfl_dt(1.0)
    dt = (dt0 * TWO_PI / A).contiguous()                      # CFL rule with the trajectory's own speed
    out = {}
    ms = timeit(lambda: out.update(s.fwd_adj(u0, A, dt, S, want_uT=False)))
    sums = s.reduce_indicators(out["eta"], out["J"])
    ups = 2 * 5 * S * K * B
    fl = 0.5 * ((2 * 25 + 60 + 6) + (2 * 25 + 60 + 8))
    peak, _ = s.measure_dfma_peak(0.5)
    return dict(workload="config 4: parameter sweep N=4 K=64 S=100, 1024 speeds x %d pulse widths per GPU" % (B // n1),
                metric="DG element-stage updates/s (fwd+adjoint)", value=ups / (ms * 1e-3), ms=ms, B=B,
                frac_fp64_peak=fl * ups / (ms * 1e-3) / 1e12 / peak, norms=dict(sum_abs_eta=float(sums[K]), max_abs_eta=float(sums[K + 2]), sum_J=float(sums[K + 3])),
                plan=s.plan(B))


def forward(B=32768):
    """The reference's own workload (utils/One_code.mlx): the forward LSERK4 march alone, config 2 mesh."""
    N, K, S = 8, 1024, 100
    s = pkg.AdvecDG1D(N, K, domain=(0.0, TWO_PI), alpha=0.0, bc="periodic")
    g = torch.Generator(device=dev); g.manual_seed(1234)
    x = torch.tensor(s.g.x, device=dev)[None]
    u0 = torch.zeros((B, N + 1, K), dtype=torch.float64, device=dev)
    for m in range(1, 5):
        A = torch.randn((B, 1, 1), dtype=torch.float64, device=dev, generator=g) / m
        ph = torch.rand((B, 1, 1), dtype=torch.float64, device=dev, generator=g) * TWO_PI
        u0 += A * torch.sin(m * x + ph)
    dt, _ = s.cfl_dt(1.0)
    ms = timeit(lambda: s.forward(u0, TWO_PI, dt, S), warm=2, reps=3)
    ups = 5 * S * K * B
    fl = 2 * 81 + 12 * 9 + 6                                   # forward stage, Np = 9
    peak, _ = s.measure_dfma_peak(0.5)
    return dict(workload="forward LSERK4 march only (One_code.mlx loop), N=8 K=1024 B=%d S=%d, periodic, upwind" % (B, S),
                metric="DG element-stage updates/s (forward)", value=ups / (ms * 1e-3), ms=ms,
                frac_fp64_peak=fl * ups / (ms * 1e-3) / 1e12 / peak, plan=s.plan(B, fused=False))


def long_march(B=2368):
    """A march too long for the one-pass kernel's residual ring: two-level checkpointing."""
    N, K, S = 8, 1024, 20000
    s = pkg.AdvecDG1D(N, K, domain=(0.0, TWO_PI), alpha=0.0, bc="periodic")
    x = torch.tensor(s.g.x, device=dev)[None]
    g = torch.Generator(device=dev); g.manual_seed(1234)
    ph = torch.rand((B, 1, 1), dtype=torch.float64, device=dev, generator=g) * TWO_PI
    u0 = torch.sin(x + ph).contiguous()
    dt, _ = s.cfl_dt(1.0)
    refused = ""
    try:
        s.fwd_adj(u0, TWO_PI, dt, S, want_uT=False)
    except pkg.DgadjError as e:
        refused = str(e)
    W = int(math.ceil(math.sqrt(S)))
    out = {}
    ms = timeit(lambda: out.update(s.fwd_adj(u0, TWO_PI, dt, S, want_uT=False, window=W)), warm=0, reps=1)
    ups = 2 * 5 * S * K * B
    return dict(workload="long march N=8 K=1024 B=%d S=%d (T = %.3f): windows of %d steps, two-level checkpointing"
                         % (B, S, S * dt, W), metric="DG element-stage updates/s (fwd+adjoint)", value=ups / (ms * 1e-3), ms=ms,
                one_pass_kernel=refused or "ran", sum_abs_eta=float(out["eta"].abs().sum()), J_mean=float(out["J"].mean()))


def burgers(B=16384):
    N, K = 4, 256
    s = pkg.BurgersDG1D(N, K, domain=(-1.0, 1.0), bc="periodic")
    g = torch.Generator(device=dev); g.manual_seed(1235)
    c = torch.rand((B, 1, 1), dtype=torch.float64, device=dev, generator=g) - 0.5
    A = torch.rand((B, 1, 1), dtype=torch.float64, device=dev, generator=g) + 0.5
    ph = torch.rand((B, 1, 1), dtype=torch.float64, device=dev, generator=g) * TWO_PI
    x = torch.tensor(s.g.x, device=dev)[None]
    u0 = (c + A * torch.sin(math.pi * x + ph)).contiguous()
    dt = s.stable_dt(2.0)
    S = int(os.environ.get("SEC_S", 200))
    ms = timeit(lambda: s.forward(u0, dt, S, checkpoints=True), warm=2, reps=3)
    out = s.forward(u0, dt, S, checkpoints=True)
    ups = 5 * S * K * B
    r1 = dict(workload="config 3: Burgers + SlopeLimitN every stage, N=4 K=256 B=%d S=%d, forward with all checkpoints "
                       "(states, limiter flags/branches, argmax, wave speeds)" % (B, S),
              metric="DG element-stage updates/s (forward, limited)", value=ups / (ms * 1e-3), ms=ms,
              limited_fraction=float(((out["lim"].int() & 31) != 0).double().mean()), T=S * dt)
    ms2 = timeit(lambda: s.adjoint(out), warm=1, reps=3)
    r2 = dict(workload="config 3: discrete adjoint of that march (frozen limiter / minmod / argmax branches), dJ/du0",
              metric="DG element-stage updates/s (adjoint incl. stage-state recompute)", value=ups / (ms2 * 1e-3), ms=ms2)
    return [r1, r2]


def tdg_fd(B=4096):
    rng = np.random.default_rng(0)
    y0 = torch.tensor(rng.uniform(-3, 3, B), device=dev)
    times = np.sort(np.concatenate(([0.0, 2.0], rng.uniform(0.05, 1.95, 30))))
    Ks = times.size - 1
    Ns = np.ones(Ks, dtype=int)
    s = pkg.TimeDG()
    res = {}

    def run():
        t1, y1, its = s.dg_march(Ns, Ks, times, y0)
        t2, v, err = s.adj_march(Ns + 1, Ks, times, y1, t1)
        res.update(its=its)
    ms = timeit(run)
    its = res["its"].double()
    nq = 31
    r1 = dict(workload="config 5: time-DG march (Newton, u'=sin u) + adjoint + indicator, n=1, Ks=%d, B=%d" % (Ks, B),
              metric="element-solves/s (forward Newton solve or adjoint solve+indicator)", value=2 * Ks * B / (ms * 1e-3), ms=ms,
              mean_newton_its=float(its.mean()), sincos_per_s=float((its.sum() * nq + B * Ks * 5) / (ms * 1e-3)),
              note="ms includes the host-side per-element constant setup (numpy) of both calls")
    f = pkg.FDAdjoint()
    ms = timeit(lambda: f.solve(y0, np.diff(times), want=("err_steps", "ref_idx")))
    n = Ks
    r2 = dict(workload="config 5: FD path forwardSolve+adjSolve+errEst+windows+argmax, n=%d steps, ref_factor 4, B=%d" % (n, B),
              metric="fine-step updates/s (fwd step, adjoint step, residual step each count 1)", value=(n + 2 * 4 * n) * B / (ms * 1e-3), ms=ms)
    Bbig = 1 << 20
    y0b = torch.tensor(rng.uniform(-3, 3, Bbig), device=dev)
    ms = timeit(lambda: f.solve(y0b, np.diff(times), want=("err_steps", "ref_idx")))
    r3 = dict(r2, workload=r2["workload"].replace("B=%d" % B, "B=%d" % Bbig), value=(n + 2 * 4 * n) * Bbig / (ms * 1e-3), ms=ms)
    return [r1, r2, r3]


if __name__ == "__main__":
    which = sys.argv[1:] or ["forward", "sweep", "burgers", "tdg_fd"]
    for w in which:
        kw = dict(B=int(os.environ["SEC_B"])) if "SEC_B" in os.environ else {}
        r = dict(forward=forward, long_march=long_march, sweep=sweep, burgers=burgers, tdg_fd=tdg_fd)[w](**kw)
        for line in (r if isinstance(r, list) else [r]):
            print(json.dumps(line), flush=True)
