"""Oracle: inviscid Burgers DG march with the reference's SlopeLimitN fused after every RK
stage (BASELINE config 3).  PARITY UNPINNED: the reference ships the limiter family
(utils/SlopeLimitN.m ...) but no Burgers right-hand side; this one is build-specified
 in the form of the toolkit the reference's utils/ come from:

    f = u^2/2,  C = max|u| over the mesh (per trajectory, per stage)
    flux = nx (f^- - f^+)/2 - C/2 (u^- - u^+)        (local Lax-Friedrichs, strong form)
    rhs  = -rx o (Dr f) + LIFT (Fscale o flux)

Time stepping: the LSERK4 loop of utils/One_code.mlx with `u = SlopeLimitN(u)` after every
stage update (and once on the initial state).  TEST INFRASTRUCTURE, see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np

from . import limiter
from . import operators as ops

BC_PERIODIC = "periodic"
BC_FREE = "free"        # ghost state = interior trace (zero jump) at both ends


def BurgersRHS1D(u, g, bc=BC_PERIODIC):
    Np = g.Np
    um0, um1 = u[..., 0, :], u[..., Np - 1, :]
    up0 = np.roll(um1, 1, axis=-1)
    up1 = np.roll(um0, -1, axis=-1)
    if bc != BC_PERIODIC:
        up0 = up0.copy(); up1 = up1.copy()
        up0[..., 0] = um0[..., 0]
        up1[..., -1] = um1[..., -1]
    maxvel = np.max(np.abs(u), axis=(-2, -1))
    mv = maxvel[..., None] if np.ndim(maxvel) else maxvel
    flux0 = -1.0 * ((um0 ** 2 - up0 ** 2) / 2.0) / 2.0 - mv / 2.0 * (um0 - up0)     # nx = -1
    flux1 = +1.0 * ((um1 ** 2 - up1 ** 2) / 2.0) / 2.0 - mv / 2.0 * (um1 - up1)     # nx = +1
    flux = np.stack([flux0, flux1], axis=-2)
    return -g.rx * (g.Dr @ (u ** 2 / 2.0)) + g.LIFT @ (g.Fscale * flux), maxvel


def burgers_march(u0, g, dt, nsteps, bc=BC_PERIODIC, limit=True, history=False):
    """Returns (uT, hist, flags, maxvel): hist[n] = u^n (n = 0..S); flags[n, s] = cells limited
    after stage s of step n; maxvel[n, s]."""
    periodic = bc == BC_PERIODIC
    u = np.array(u0, dtype=float, copy=True)
    if limit:
        u = limiter.SlopeLimitN(u, g, periodic)
    resu = np.zeros_like(u)
    hist = [u.copy()] if history else None
    flags, mvs = [], []
    for _ in range(nsteps):
        fl, mv = [], []
        for s in range(5):
            rhs, maxvel = BurgersRHS1D(u, g, bc)
            resu = ops.rk4a[s] * resu + dt * rhs
            u = u + ops.rk4b[s] * resu
            if limit:
                u, ids = limiter.SlopeLimitN(u, g, periodic, return_flags=True)
            else:
                ids = np.zeros(u.shape[:-2] + (u.shape[-1],), dtype=bool)
            fl.append(ids); mv.append(maxvel)
        flags.append(np.stack(fl, axis=0)); mvs.append(np.stack(mv, axis=0))
        if history:
            hist.append(u.copy())
    return u, (np.stack(hist, axis=0) if history else None), np.stack(flags, axis=0), np.stack(mvs, axis=0)
