#!/usr/bin/env python
"""BASELINE config 5 as a report: adjoint-driven refinement of the DG-in-time
path (matlab/MAIN.m loop) against the finite-difference path (python/Main_finite_difference.py
loop) on the same B = 4096 initial conditions u0 ~ U(-3, 3), u' = sin(u) on [0, 2], shared mesh,
batch-mean indicator, 30 argmax refinements: batch-mean |J_h - J_exact| of J = int_0^2 u dt and the
indicator totals against the number of elements.  J_exact from the closed form of
python/factory.py:130-131, u(t) = 2 atan2(sin(u0/2) e^t, cos(u0/2)).  One table on stdout."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import dgadj_loader

pkg = dgadj_loader.load_package()
from adjoint_ode_adaptivity_b200.galerkin import BaseGalerkin1D

B, T, ITERS = 4096, 2.0, 30
rng = np.random.default_rng(0)
u0 = rng.uniform(-3, 3, B)
d_u0 = torch.tensor(u0, device="cuda")

xq, wq = np.polynomial.legendre.leggauss(200)
tq = 0.5 * T * (xq + 1.0)
J_exact = (0.5 * T * wq * 2.0 * np.arctan2(np.sin(u0[:, None] / 2) * np.exp(tq), np.cos(u0[:, None] / 2))).sum(1)

QUIRKS = "--no-quirks" not in sys.argv      # --no-quirks: SURVEY quirk C-3 off (TimeDG(quirks=False))
torch.cuda.synchronize(); t0 = time.perf_counter()
hist_dg = pkg.adapt_tdg(d_u0, tspan=(0.0, T), Ks=2, n=1, iters=ITERS, quirks=QUIRKS)
torch.cuda.synchronize(); t_dg = time.perf_counter() - t0
t0 = time.perf_counter()
hist_fd = pkg.adapt_fd(d_u0, tspan=(0.0, T), n_steps=2, iters=ITERS, functional="int_u")
torch.cuda.synchronize(); t_fd = time.perf_counter() - t0

s = pkg.TimeDG()
f = pkg.FDAdjoint(functional="int_u")
print("config 5: B=%d ICs, u'=sin u, T=%g, %d refinements, DG-in-time %s; loop wall time DG-in-time %.3f s, FD %.3f s"
      % (B, T, ITERS, "bug for bug" if QUIRKS else "with quirk C-3 off", t_dg, t_fd))
print("%4s | %9s %14s %14s | %9s %14s %14s" % ("it", "DG elems", "mean|J_h-J|", "sum mean|err|", "FD steps", "mean|J_h-J|", "sum mean err"))
for it in (0, 5, 10, 20, 30):
    times = hist_dg[it]["times"]
    Ks = times.size - 1
    t1, y1, _ = s.dg_march(np.ones(Ks, dtype=int), Ks, times, d_u0)
    w = np.stack([BaseGalerkin1D(n=1, k=1, domain=(times[k], times[k + 1])).quad_weights()[:, 0] for k in range(Ks)])
    J_dg = (y1.cpu().numpy() * w[None]).sum((1, 2))
    tf = hist_fd[it]["times"]
    u = f.solve(d_u0, np.diff(tf), want=("u",))["u"].cpu().numpy()
    J_fd = (u[:, :-1] * np.diff(tf)[None]).sum(1)                       # J = sum u_n dt_n (getK, :153-155)
    print("%4d | %9d %14.3e %14.3e | %9d %14.3e %14.3e" % (it, Ks, np.abs(J_dg - J_exact).mean(), hist_dg[it]["err_total"],
                                                            tf.size - 1, np.abs(J_fd - J_exact).mean(), hist_fd[it]["err_total"]))
