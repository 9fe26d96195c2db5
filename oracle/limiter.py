"""Oracle: the slope-limiter family of the reference, utils/{minmod,minmodB,SlopeLimitLin,
SlopeLimitN,SlopeLimit1}.m, restated in NumPy and batched over leading axes (TEST
INFRASTRUCTURE, see oracle/__init__.py).  Arrays are (..., Np, K).

The reference never calls these routines, so there is no reference
output to pin them against beyond the arithmetic itself; SURVEY App. B.5 holds a
known-answer case.  Quirk C-16 (ghost cell averages copy the end cells, SlopeLimitN.m:18) is
the default; `periodic=True` wraps them instead (build option for periodic problems).
"""
from __future__ import annotations

import numpy as np


def minmod(v):
    """utils/minmod.m:6-12.  v: (m, ...) -> (...): s*min|v| where all m signs agree, else 0."""
    m = v.shape[0]
    s = np.sum(np.sign(v), axis=0) / m
    out = np.zeros(v.shape[1:])
    ids = np.abs(s) == 1
    out[ids] = (s * np.min(np.abs(v), axis=0))[ids]
    return out


def minmodB(v, M, h):
    """utils/minmodB.m:6-11 (TVB modification)."""
    mfunc = v[0].copy()
    ids = np.abs(mfunc) > M * h ** 2
    if np.any(ids):
        mfunc[ids] = minmod(v)[ids]
    return mfunc


def _neighbour_averages(v, periodic):
    if periodic:
        return np.roll(v, 1, axis=-1), np.roll(v, -1, axis=-1)
    vkm1 = np.concatenate((v[..., :1], v[..., :-1]), axis=-1)       # [v(1), v(1:K-1)]
    vkp1 = np.concatenate((v[..., 1:], v[..., -1:]), axis=-1)       # [v(2:K), v(K)]
    return vkm1, vkp1


def SlopeLimitLin(ul, xl, vm1, v0, vp1, g, M=0.0):
    """utils/SlopeLimitLin.m:10-18.  M > 0: the TVB form -- minmodB(., M, h) (utils/minmodB.m:6-11) in
    place of minmod at :18, the way minmodB is meant to be used; M = 0 is the reference's call."""
    Np = g.Np
    h = xl[Np - 1, :] - xl[0, :]
    x0 = xl[0, :] + h / 2
    ux = (2.0 / h) * (g.Dr @ ul)
    args = np.stack([ux[..., 0, :], (vp1 - v0) / h, (v0 - vm1) / h], axis=0)
    slope = minmodB(args, M, h) if M > 0.0 else minmod(args)
    return v0[..., None, :] + (xl - x0) * slope[..., None, :]


def cell_averages(u, g):
    """uh = invV*u; uh(2:Np,:) = 0; uavg = V*uh; v = uavg(1,:)   (SlopeLimitN.m:9)."""
    uh0 = g.invV[0:1, :] @ u                                        # (..., 1, K)
    return (g.V[0, 0] * uh0)[..., 0, :]


def SlopeLimitN(u, g, periodic=False, return_flags=False, M=0.0):
    """utils/SlopeLimitN.m:9-32 (Pi^N: detect, then limit the flagged cells)."""
    eps0 = 1.0e-8
    v = cell_averages(u, g)
    ue1, ue2 = u[..., 0, :], u[..., -1, :]
    vkm1, vkp1 = _neighbour_averages(v, periodic)
    ve1 = v - minmod(np.stack([v - ue1, v - vkm1, vkp1 - v], axis=0))
    ve2 = v + minmod(np.stack([ue2 - v, v - vkm1, vkp1 - v], axis=0))
    ids = (np.abs(ve1 - ue1) > eps0) | (np.abs(ve2 - ue2) > eps0)
    ulimit = u.copy()
    if np.any(ids):
        uhl = g.invV @ u
        uhl[..., 2:, :] = 0.0
        ul = g.V @ uhl
        lim = SlopeLimitLin(ul, g.x, vkm1, v, vkp1, g, M)
        mask = np.broadcast_to(ids[..., None, :], u.shape)
        ulimit[mask] = lim[mask]
    return (ulimit, ids) if return_flags else ulimit


def SlopeLimit1(u, g, periodic=False, M=0.0):
    """utils/SlopeLimit1.m:10-22 (Pi^1: limit every cell)."""
    uh = g.invV @ u
    ul = uh.copy()
    ul[..., 2:, :] = 0.0
    ul = g.V @ ul
    v = cell_averages(u, g)
    vkm1, vkp1 = _neighbour_averages(v, periodic)
    return SlopeLimitLin(ul, g.x, vkm1, v, vkp1, g, M)
