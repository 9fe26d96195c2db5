"""The reference's MATLAB entry points under their own names, as thin wrappers over the device
handles (argument order and meaning follow the .m files; batching adds a leading axis).

    rhsu        = AdvecRHS1D(solver, u, timelocal, a)            utils/AdvecRHS1D.m:1
    ulimit      = SlopeLimitN(burgers_solver, u)                  utils/SlopeLimitN.m:1
    ulimit      = SlopeLimit1(burgers_solver, u)                  utils/SlopeLimit1.m:1
    [t, y]      = dg_march(tdg, Ns, Ks, times, y0)                matlab/dg_march.m:1
    [t, v, err] = adj_march(tdg, Ns, Ks, times, y1, t1)           matlab/adj_march.m:1  (primal passed
                                                                  explicitly instead of the globals y1, t1)
    [t, v, err] = adj_rec(tdg, Ns, Ks, times, y1, t1)             matlab/adj_rec.m:1
    [err, res]  = err_contribution(tdg, Ks, Ns, uh, t1)           matlab/err_contribution.m:1
    [t, y]      = fwd_euler_march(y0, times, ode)                 matlab/fwd_euler_march.m:1 (a broken stub in
                                                                  the reference; semantics of forwardSolve,
                                                                  python/Main_finite_difference.py:34-51)
"""
from __future__ import annotations

import numpy as np

from .fd import FDAdjoint


def AdvecRHS1D(solver, u, timelocal, a):
    return solver.rhs(u, timelocal, a)


def SlopeLimitN(burgers_solver, u):
    return burgers_solver.slope_limit(u)


def SlopeLimit1(burgers_solver, u):
    return burgers_solver.slope_limit(u, kind="1")


def dg_march(tdg, Ns, Ks, times, y0, x_true=None, u_true=None):
    t, y, _ = tdg.dg_march(Ns, Ks, times, y0, x_true, u_true)
    return t, y


def adj_march(tdg, Ns, Ks, times, y1, t1):
    return tdg.adj_march(Ns, Ks, times, y1, t1)


def adj_rec(tdg, Ns, Ks, times, y1, t1):
    return tdg.adj_rec(Ns, Ks, times, y1, t1)


def err_contribution(tdg, Ks, Ns, uh, t1):
    return tdg.err_contribution(Ks, Ns, uh, t1), [None] * Ks


def fwd_euler_march(y0, times, ode="sin", device=0):
    s = FDAdjoint(ode=ode, device=device)
    try:
        y = s.forwardSolve(np.diff(np.asarray(times, dtype=np.float64)), y0)
    finally:
        s.close()
    return np.asarray(times, dtype=np.float64), y
