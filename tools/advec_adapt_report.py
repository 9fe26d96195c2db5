#!/usr/bin/env python
"""Goal-oriented h-refinement of the DG-in-space advection march (the headline path used the way
the reference uses its ODE loop, matlab/MAIN.m:29-166): a batch of Gaussian pulses, a = 1,
periodic on [0, 2 pi], functional J = int psi(x) u(x, T) dx with a weight psi centred downstream; per
iteration the fused forward + adjoint + indicator kernel, the batch-mean indicator, the two
elements with the largest mean |eta| split.  Against the closed form J_exact = int psi(x) u0(x - aT) dx
(evaluated by quadrature): batch-mean |J_h - J_exact|, the indicator totals, and -- the property the
indicator is built for -- how well sum_k eta_k = J_h - J_(order N+1) predicts the error J_h - J_exact."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import dgadj_loader

pkg = dgadj_loader.load_package()

B, N, a, T = 256, 4, 1.0, 1.5
rng = np.random.default_rng(3)
centres = rng.uniform(1.5, 2.5, B)
widths = rng.uniform(0.15, 0.3, B)
psi = lambda x: 0.5 * (1.0 + np.cos(x - 4.0))      # smooth and resolved by every mesh: no quadrature error in J_h


def u0_np(x):
    return np.exp(-((x[None] - centres[:, None, None]) / widths[:, None, None]) ** 2)


xq, wq = np.polynomial.legendre.leggauss(4000)
xx = math.pi * (xq + 1.0)
sh = np.mod(xx[None] - a * T, 2 * math.pi)
J_exact = (math.pi * wq * psi(xx)[None] * np.exp(-((sh - centres[:, None]) / widths[:, None]) ** 2)).sum(1)

u0_fn = lambda x: torch.tensor(u0_np(x), device="cuda")
hist = pkg.adapt_advec(u0_fn, N, np.linspace(0.0, 2 * math.pi, 9), a, T, iters=14, topk=2, bc="periodic", alpha=0.0, psi=psi)
print("goal-oriented h-refinement, N=%d, B=%d pulses, J = int psi u(T): mean|J_h - J|, indicator totals" % (N, B))
print("%3s %4s %6s %13s %15s %24s" % ("it", "K", "S", "mean|J_h-J|", "sum mean|eta|", "mean|sum eta-(J_h-J)|"))
for h in hist:
    J_h = h["J"].cpu().numpy()
    est = h["estimate"].cpu().numpy()
    print("%3d %4d %6d %13.3e %15.3e %24.3e" % (h["it"], h["K"], h["S"], np.abs(J_h - J_exact).mean(), h["eta_total"],
                                                   np.abs(est - (J_h - J_exact)).mean()))
