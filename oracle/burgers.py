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


def burgers_march(u0, g, dt, nsteps, bc=BC_PERIODIC, limit=True, history=False, tvb_M=0.0):
    """Returns (uT, hist, flags, maxvel): hist[n] = u^n (n = 0..S); flags[n, s] = cells limited
    after stage s of step n; maxvel[n, s].  limit: True / "N" = SlopeLimitN, "1" = SlopeLimit1 (every
    cell), False = none; tvb_M: the M of minmodB."""
    periodic = bc == BC_PERIODIC
    u = np.array(u0, dtype=float, copy=True)

    def apply(u):
        if limit == "1":
            return limiter.SlopeLimit1(u, g, periodic, tvb_M), np.ones(u.shape[:-2] + (u.shape[-1],), dtype=bool)
        return limiter.SlopeLimitN(u, g, periodic, return_flags=True, M=tvb_M)
    if limit:
        u = apply(u)[0]
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
                u, ids = apply(u)
            else:
                ids = np.zeros(u.shape[:-2] + (u.shape[-1],), dtype=bool)
            fl.append(ids); mv.append(maxvel)
        flags.append(np.stack(fl, axis=0)); mvs.append(np.stack(mv, axis=0))
        if history:
            hist.append(u.copy())
    return u, (np.stack(hist, axis=0) if history else None), np.stack(flags, axis=0), np.stack(mvs, axis=0)


# ----------------------------------------------------------------------------------
# Discrete adjoint of the limited Burgers march (build-specified; PARITY UNPINNED).
# The limiter and max|u| are piecewise linear / piecewise smooth: the adjoint transposes the
# linearisation on the branches the forward run took (recorded per stage: which cells were
# limited, which minmod argument won, where max|u| sat) -- SURVEY section 7, hard part 5.
# ----------------------------------------------------------------------------------
def _minmod_branch(a, b, c):
    """minmod of three rows plus the index (1..3) of the winning argument, 0 where the result
    is 0 (signs differ).  Ties go to the lowest index."""
    v = np.stack([a, b, c], axis=0)
    s = np.sum(np.sign(v), axis=0) / 3.0
    ok = np.abs(s) == 1
    av = np.abs(v)
    win = np.argmin(av, axis=0)                       # first minimum
    val = np.where(ok, s * np.min(av, axis=0), 0.0)
    return val, np.where(ok, win + 1, 0)


def limit_with_branches(u, g, periodic):
    """SlopeLimitN (same arithmetic as oracle/limiter.py) returning also flags and branches."""
    eps0 = 1.0e-8
    v = limiter.cell_averages(u, g)
    ue1, ue2 = u[..., 0, :], u[..., -1, :]
    vkm1, vkp1 = limiter._neighbour_averages(v, periodic)
    ve1 = v - limiter.minmod(np.stack([v - ue1, v - vkm1, vkp1 - v], axis=0))
    ve2 = v + limiter.minmod(np.stack([ue2 - v, v - vkm1, vkp1 - v], axis=0))
    ids = (np.abs(ve1 - ue1) > eps0) | (np.abs(ve2 - ue2) > eps0)
    uhl = g.invV @ u
    uhl[..., 2:, :] = 0.0
    ul = g.V @ uhl
    Np = g.Np
    h = g.x[Np - 1, :] - g.x[0, :]
    x0 = g.x[0, :] + h / 2
    ux = (2.0 / h) * (g.Dr @ ul)
    slope, br = _minmod_branch(ux[..., 0, :], (vkp1 - v) / h, (v - vkm1) / h)
    lim = v[..., None, :] + (g.x - x0) * slope[..., None, :]
    out = np.where(ids[..., None, :], lim, u)
    return out, ids, np.where(ids, br, 0)


def limiter_weights(g):
    """aw (cell average), sl (slope of the linear part at node 1), xc = x - x0, h."""
    Np = g.Np
    aw = g.V[0, 0] * g.invV[0, :]
    sl = g.Dr[0, :] @ (g.V[:, :2] @ g.invV[:2, :])
    h = g.x[Np - 1, :] - g.x[0, :]
    xc = g.x - (g.x[0, :] + h / 2)
    return aw, sl, xc, h


def limiter_T(lam, ids, br, g, periodic):
    """Transpose of the frozen-branch limiter:  u' = v + xc*s on limited cells, identity else."""
    aw, sl, xc, h = limiter_weights(g)
    a = np.sum(lam, axis=-2)                                   # multiplies v
    c = np.sum(xc * lam, axis=-2)                              # multiplies the slope
    lv = np.where(ids, a, 0.0)                                 # d/dv_k
    ch = c / h
    own = np.where(br == 2, -ch, 0.0) + np.where(br == 3, ch, 0.0)
    to_right = np.where(br == 2, ch, 0.0)                      # adds to v_{k+1}
    to_left = np.where(br == 3, -ch, 0.0)                      # adds to v_{k-1}
    lv = lv + own
    if periodic:
        lv = lv + np.roll(to_right, 1, axis=-1) + np.roll(to_left, -1, axis=-1)
    else:
        r = np.zeros_like(lv); r[..., 1:] = to_right[..., :-1]
        l = np.zeros_like(lv); l[..., :-1] = to_left[..., 1:]
        # the end cells see a copied ghost average: vkm1(1) = v(1), vkp1(K) = v(K)
        r[..., -1] += to_right[..., -1]
        l[..., 0] += to_left[..., 0]
        lv = lv + r + l
    out = np.where(ids[..., None, :], 0.0, lam)
    out = out + aw[:, None] * lv[..., None, :]
    out = out + np.where((br == 1)[..., None, :], sl[:, None] * ((2.0 / h) * c)[..., None, :], 0.0)
    return out


def burgers_record(u0, g, dt, nsteps, bc=BC_PERIODIC):
    """Forward march keeping everything the adjoint needs: stage input states, limiter flags /
    branches, location and sign of max|u| per stage."""
    periodic = bc == BC_PERIODIC
    u = np.array(u0, dtype=float, copy=True)
    u, ids0, br0 = limit_with_branches(u, g, periodic)
    rec = dict(ids0=ids0, br0=br0, stages=[])
    resu = np.zeros_like(u)
    for _ in range(nsteps):
        for s in range(5):
            flat = np.abs(u).reshape(u.shape[:-2] + (-1,))
            am = np.argmax(flat, axis=-1)                      # first maximum, row-major (i, k)
            sg = np.sign(np.take_along_axis(u.reshape(flat.shape), am[..., None], axis=-1))[..., 0]
            rhs, maxvel = BurgersRHS1D(u, g, bc)
            resu = ops.rk4a[s] * resu + dt * rhs
            ut = u + ops.rk4b[s] * resu
            un, ids, br = limit_with_branches(ut, g, periodic)
            rec["stages"].append(dict(u=u, am=am, sg=sg, maxvel=maxvel, ids=ids, br=br, s=s))
            u = un
    rec["uT"] = u
    return rec


def BurgersRHS1D_T(lam, st, g, bc):
    """(dR/du)^T lam at the stage state st['u'], including the dependence through C = max|u|."""
    u, maxvel = st["u"], st["maxvel"]
    Np = g.Np
    periodic = bc == BC_PERIODIC
    mv = maxvel[..., None] if np.ndim(maxvel) else maxvel
    um0, um1 = u[..., 0, :], u[..., Np - 1, :]
    up0 = np.roll(um1, 1, axis=-1)
    up1 = np.roll(um0, -1, axis=-1)
    if not periodic:
        up0 = up0.copy(); up1 = up1.copy()
        up0[..., 0] = um0[..., 0]
        up1[..., -1] = um1[..., -1]
    G = np.swapaxes(g.LIFT, 0, 1) @ lam                         # (..., 2, K)
    G0 = G[..., 0, :] * g.Fscale[0, :]
    G1 = G[..., 1, :] * g.Fscale[1, :]
    # flux0 = -(um0^2 - up0^2)/4 - C/2 (um0 - up0);  flux1 = +(um1^2 - up1^2)/4 - C/2 (um1 - up1)
    d0_m, d0_p = (-um0 / 2.0 - mv / 2.0) * G0, (up0 / 2.0 + mv / 2.0) * G0
    d1_m, d1_p = (um1 / 2.0 - mv / 2.0) * G1, (-up1 / 2.0 + mv / 2.0) * G1
    out = u * (g.Dr.T @ (-g.rx * lam))                          # volume: -rx Dr (u^2/2)
    out[..., 0, :] += d0_m
    out[..., Np - 1, :] += d1_m
    # neighbour contributions: up0[k] = um1[k-1], up1[k] = um0[k+1]
    if periodic:
        out[..., Np - 1, :] += np.roll(d0_p, -1, axis=-1)
        out[..., 0, :] += np.roll(d1_p, 1, axis=-1)
    else:
        out[..., Np - 1, :-1] += d0_p[..., 1:]
        out[..., 0, 1:] += d1_p[..., :-1]
        out[..., 0, 0] += d0_p[..., 0]                          # ghost = own trace
        out[..., Np - 1, -1] += d1_p[..., -1]
    # through C: dflux/dC = -(u^- - u^+)/2
    gam = np.sum(G0 * (-(um0 - up0) / 2.0) + G1 * (-(um1 - up1) / 2.0), axis=-1)
    flat = out.reshape(out.shape[:-2] + (-1,))
    np.put_along_axis(flat, st["am"][..., None],
                      np.take_along_axis(flat, st["am"][..., None], axis=-1) + (gam * st["sg"])[..., None], axis=-1)
    return flat.reshape(out.shape)


def burgers_adjoint(rec, g, dt, lamT, bc=BC_PERIODIC):
    """Reverse sweep: lam0 = dJ/du0 for J with dJ/du^S = lamT (frozen limiter / max branches)."""
    periodic = bc == BC_PERIODIC
    lu = np.array(lamT, dtype=float, copy=True)
    lk = np.zeros_like(lu)
    for st in reversed(rec["stages"]):
        s = st["s"]
        lu = limiter_T(lu, st["ids"], st["br"], g, periodic)   # through u' = L(u + rkb res')
        lk = lk + ops.rk4b[s] * lu
        lu = lu + dt * BurgersRHS1D_T(lk, st, g, bc)
        lk = ops.rk4a[s] * lk
    return limiter_T(lu, rec["ids0"], rec["br0"], g, periodic)


# ----------------------------------------------------------------------------------
# Per-element indicator of the limited Burgers march (build-specified; PARITY UNPINNED), the nonlinear
# form of SURVEY App. E.5 and of the reference's estimators (matlab/MAIN.m:34 solves the adjoint at
# order Ns+1; matlab/adj_march.m:103-117: err(k) = v_k' * residual; errEst,
# python/Main_finite_difference.py:79-94: res[n] = u_f[n] - fwdUpdate(u_f, dt_f, n), err = res*v):
#   rho^n   = P u^{n+1} - Phi_f(P u^n)          enriched-space (order N+1) one-step residual of the coarse
#                                               march; Phi_f = the same limited LSERK4 step at order N+1;
#                                               P = nodal prolongation V_f(:,1:Np) inv(V_c)
#   lam_f^n = Phi_f'(P u^n)^T lam_f^{n+1}       enriched-space discrete adjoint, linearised at the prolonged
#                                               coarse states on the frozen limiter / minmod / argmax
#                                               branches of those steps;  lam_f^S = jw_f
#   eta_k   = sum_n lam_f^{n+1}_k . rho^n_k
# so that sum_k eta_k = J_f(P u_c^S) - J_f(u_f^S) to first order in the residuals (u_f = the enriched
# march of P u_c^0) -- an identity for a linear problem (oracle/advec.py).  The coarse-space adjoint
# (burgers_adjoint above: the exact gradient of the coarse march) does NOT give a usable estimate: it
# differs from lam_f most in the modes the residual lives in (tests/test_oracle_golden.py).
# ----------------------------------------------------------------------------------
def burgers_step_record(u, g, dt, bc=BC_PERIODIC):
    """One limited LSERK4 step from u (no limiter pass on u itself) keeping what its transpose
    needs: per stage the input state, limiter flags / branches, location and sign of max|u|."""
    periodic = bc == BC_PERIODIC
    u = np.array(u, dtype=float, copy=True)
    resu = np.zeros_like(u)
    stages = []
    for s in range(5):
        flat = np.abs(u).reshape(u.shape[:-2] + (-1,))
        am = np.argmax(flat, axis=-1)
        sg = np.sign(np.take_along_axis(u.reshape(flat.shape), am[..., None], axis=-1))[..., 0]
        rhs, maxvel = BurgersRHS1D(u, g, bc)
        resu = ops.rk4a[s] * resu + dt * rhs
        ut = u + ops.rk4b[s] * resu
        un, ids, br = limit_with_branches(ut, g, periodic)
        stages.append(dict(u=u, am=am, sg=sg, maxvel=maxvel, ids=ids, br=br, s=s))
        u = un
    return u, stages


def burgers_step_T(lu, stages, g, dt, bc=BC_PERIODIC):
    """Transpose of the linearised step recorded by burgers_step_record (the RK residual starts and
    ends a step with weight zero: rk4a[0] = 0)."""
    periodic = bc == BC_PERIODIC
    lk = np.zeros_like(lu)
    for st in reversed(stages):
        s = st["s"]
        lu = limiter_T(lu, st["ids"], st["br"], g, periodic)
        lk = lk + ops.rk4b[s] * lu
        lu = lu + dt * BurgersRHS1D_T(lk, st, g, bc)
        lk = ops.rk4a[s] * lk
    return lu


def prolongation(gc, gf):
    """P[NpF, Np] = V_f(:, 1:Np) inv(V_c)."""
    return gf.V[:, :gc.Np] @ gc.invV


def prolong(P, u):
    """P u in the form u_1 + P (u - u_1) (u_1 = the element's first nodal value; the rows of P sum to 1):
    a cell the limiter has flattened (minmod = 0 at an extremum: all nodal values bitwise equal) stays exactly
    flat in the enriched space.  With a plain matrix product its nodes would differ by rounding, and the
    location of max|u| inside such a cell -- which the frozen-branch adjoint records -- would be decided by
    that noise instead of the lowest-index rule."""
    u1 = u[..., 0:1, :]
    return u1 + P @ (u - u1)


def burgers_fwd_adj_indicator(u0, gc, gf, dt, nsteps, jw_c, jw_f, bc=BC_PERIODIC):
    """One trajectory (u0: (Np, K)).  Returns dict(uT, J, lam0[NpF, K], eta[K], eta_scale[K] (see below), nlim,
    fine=[per step: the enriched step's record]): coarse limited march (SlopeLimitN on the initial state
    and after every stage, as burgers_march), J = sum jw_c o u(T), enriched adjoint and indicator as
    specified above; nlim = number of (cell, stage) limiter activations of the coarse march."""
    periodic = bc == BC_PERIODIC
    P = prolongation(gc, gf)
    u, ids0, _ = limit_with_branches(np.array(u0, dtype=float), gc, periodic)
    states, nlim = [u], int(np.sum(ids0)) * 0
    for _ in range(nsteps):
        u, st = burgers_step_record(u, gc, dt, bc)
        nlim += int(sum(int(np.sum(q["ids"])) for q in st))
        states.append(u)
    lam = np.array(jw_f, dtype=float, copy=True)
    K = gc.K
    eta, scale, fine = np.zeros(K), np.zeros(K), []
    for n in range(nsteps - 1, -1, -1):
        uf1, stf = burgers_step_record(prolong(P, states[n]), gf, dt, bc)
        pu1 = prolong(P, states[n + 1])
        rho = pu1 - uf1
        eta += np.sum(lam * rho, axis=0)
        # magnitude before the cancellation in rho (a difference of two O(|u|) states that agree to the
        # one-step residual): any two fp64 evaluations of eta agree to eps * scale, not to eps * |eta|
        scale += np.sum(np.abs(lam) * (np.abs(pu1) + np.abs(uf1)), axis=0)
        lam = burgers_step_T(lam, stf, gf, dt, bc)
        fine.append(stf)
    return dict(uT=states[-1], J=float(np.sum(jw_c * states[-1])), lam0=lam, eta=eta,
                eta_scale=scale + 1e-300, nlim=nlim, fine=fine[::-1], states=states)
