"""Oracle: linear-advection DG RHS, LSERK4 / forward-Euler march, discrete adjoint and the
adjoint-weighted per-element error indicator (TEST INFRASTRUCTURE, see oracle/__init__.py).

All arrays are fp64, shaped (..., Np, K) with any number of leading batch axes
(batching pattern: `vmap` over initial conditions, python/Main_variable_params.py:330-339).

Pinned by the reference: `AdvecRHS1D`, the LSERK4 loop and the CFL rule
(utils/AdvecRHS1D.m, utils/One_code.mlx).  PARITY UNPINNED (build-specified, SURVEY
App. E.5): periodic BC, upwind flux, the discrete adjoint and the indicator -- the
reference has no PDE adjoint; conventions follow matlab/adj_march.m (adjoint solved one
order higher, residual of the injected coarse solution in the enriched space) and
python/Main_finite_difference.py:79-94 (`errEst`: one-step residual times adjoint).
"""
from __future__ import annotations

import math

import numpy as np

from . import operators as ops

BC_INFLOW = "inflow"      # utils/AdvecRHS1D.m:14-16
BC_PERIODIC = "periodic"  # BASELINE config text; not in the reference

INFLOW_SIN_AT = "sin_at"      # uin = -sin(a*t)     utils/AdvecRHS1D.m:14
INFLOW_SIN_AAT = "sin_aat"    # uin = -sin(a*a*t)   utils/One_code.mlx (quirk C-2)
INFLOW_ZERO = "zero"


def _bcast(a, ndim_tail=2):
    """Per-trajectory scalar (B,) -> (B,1,1); python scalar unchanged."""
    a = np.asarray(a, dtype=float)
    return a.reshape(a.shape + (1,) * ndim_tail) if a.ndim else float(a)


def inflow_value(kind, a, t):
    if kind == INFLOW_SIN_AT:
        return -np.sin(a * t)
    if kind == INFLOW_SIN_AAT:
        return -np.sin(a * a * t)
    if kind == INFLOW_ZERO:
        return 0.0 * a
    raise ValueError(kind)


def face_coeff(a, alpha):
    """c[f] = (a*nx - (1-alpha)*|a*nx|)/2 for nx = -1, +1   (utils/AdvecRHS1D.m:11)."""
    a = np.asarray(a, dtype=float)
    c0 = (a * -1.0 - (1.0 - alpha) * np.abs(a * -1.0)) / 2.0
    c1 = (a * 1.0 - (1.0 - alpha) * np.abs(a * 1.0)) / 2.0
    return c0, c1


def AdvecRHS1D(u, timelocal, a, g, alpha=1.0, bc=BC_INFLOW, inflow=INFLOW_SIN_AT):
    """utils/AdvecRHS1D.m:8-19.  `g` = StartUp1D namespace.  a: scalar or (B,)."""
    Np = g.Np
    um0, um1 = u[..., 0, :], u[..., Np - 1, :]                   # u(vmapM)
    up0 = np.roll(um1, 1, axis=-1)                               # u(vmapP): left neighbour's right node
    up1 = np.roll(um0, -1, axis=-1)
    if bc != BC_PERIODIC:                                        # boundary faces: vmapP == vmapM
        up0 = up0.copy(); up1 = up1.copy()
        up0[..., 0] = um0[..., 0]
        up1[..., -1] = um1[..., -1]
    a_arr = np.asarray(a, dtype=float)
    ab = a_arr[..., None] if a_arr.ndim else float(a_arr)        # broadcast over K
    c0, c1 = face_coeff(ab, alpha)
    du0 = (um0 - up0) * c0                                       # :11
    du1 = (um1 - up1) * c1
    if bc == BC_INFLOW:
        uin = inflow_value(inflow, a_arr, timelocal)             # :14
        du0 = du0.copy() if du0.base is not None else du0
        du1 = du1.copy() if du1.base is not None else du1
        c00 = c0[..., 0] if np.ndim(c0) else c0
        du0[..., 0] = (u[..., 0, 0] - uin) * c00                 # :15
        du1[..., -1] = 0.0                                       # :16
    du = np.stack([du0, du1], axis=-2)                           # (..., 2, K)
    a3 = _bcast(a_arr)
    return -a3 * g.rx * (g.Dr @ u) + g.LIFT @ (g.Fscale * du)    # :19


def cfl_dt(g, FinalTime, CFL=0.75, speed=2.0 * math.pi):
    """utils/One_code.mlx time-step rule: xmin = min|x(1,:)-x(2,:)|; dt = CFL/(2*pi)*xmin;
    dt = .5*dt; Nsteps = ceil(T/dt); dt = T/Nsteps.  The divisor is literally 2*pi in the
    reference (== a there); `speed` generalises it for per-trajectory sweeps (config 4)."""
    xmin = np.min(np.abs(g.x[0, :] - g.x[1, :]))
    dt = CFL / speed * xmin
    dt = 0.5 * dt
    Nsteps = int(math.ceil(FinalTime / dt))
    return FinalTime / Nsteps, Nsteps


SCHEME_LSERK4 = (ops.rk4a, ops.rk4b, ops.rk4c)
SCHEME_EULER = (np.array([0.0]), np.array([1.0]), np.array([0.0]))   # fwd_euler_march semantics


def step(u, resu, t, dt, a, g, alpha, bc, inflow, scheme=SCHEME_LSERK4):
    """One LSERK4 step: utils/One_code.mlx inner loop `for INTRK = 1:5`."""
    A, Bc, C = scheme
    dt3 = _bcast(dt)
    dt1 = np.asarray(dt, dtype=float)
    for s in range(len(A)):
        timelocal = t + C[s] * dt1
        rhsu = AdvecRHS1D(u, timelocal, a, g, alpha, bc, inflow)
        resu = A[s] * resu + dt3 * rhsu
        u = u + Bc[s] * resu
    return u, resu


def advec_march(u0, g, a, dt, nsteps, alpha=1.0, bc=BC_INFLOW, inflow=INFLOW_SIN_AT,
                scheme=SCHEME_LSERK4, history=False, t0=0.0):
    """Forward march (utils/One_code.mlx `for tstep=1:Nsteps`).  `resu` is carried across
    steps and never re-zeroed (it is multiplied by rk4a(1)=0), as in the reference.
    Returns (uT, hist) with hist[n] = u^n (n = 0..nsteps) when history=True."""
    u = np.array(u0, dtype=float, copy=True)
    resu = np.zeros_like(u)
    hist = [u.copy()] if history else None
    time = t0 + 0.0 * np.asarray(dt, dtype=float)
    for n in range(nsteps):
        u, resu = step(u, resu, time, dt, a, g, alpha, bc, inflow, scheme)
        time = time + dt                       # `time = time+dt` accumulation, as the mlx does
        if history:
            hist.append(u.copy())
    return u, (np.stack(hist, axis=0) if history else None)


def advec_march_mlx(u0, g, a, FinalTime, alpha=1.0, inflow=INFLOW_SIN_AAT):
    """Exact restatement of the mlx driver incl. `time = time + dt` accumulation; returns
    the state of du / rhsu / resu after the last stage (the mlx 'Testing outputs')."""
    dt, Nsteps = cfl_dt(g, FinalTime)
    u = np.array(u0, dtype=float, copy=True)
    resu = np.zeros_like(u)
    time = 0.0
    rhsu = None
    for _ in range(Nsteps):
        for s in range(5):
            timelocal = time + ops.rk4c[s] * dt
            rhsu = AdvecRHS1D(u, timelocal, a, g, alpha, BC_INFLOW, inflow)
            resu = ops.rk4a[s] * resu + dt * rhsu
            u = u + ops.rk4b[s] * resu
        time = time + dt
    return dict(u=u, rhsu=rhsu, resu=resu, dt=dt, Nsteps=Nsteps, time=time)


# ----------------------------------------------------------------------------------
# Build-specified discrete adjoint -- PARITY UNPINNED
# ----------------------------------------------------------------------------------

def AdvecRHS1D_T(lam, a, g, alpha=1.0, bc=BC_INFLOW):
    """Transpose of the u-linear part of AdvecRHS1D:  L^T lam  with
    L u = -a rx o (Dr u) + LIFT (Fscale o du(u)).
        g2[f,k] = (LIFT[:,f] . lam[:,k]) * Fscale[f,k] * c[f]
        out      = -a Dr^T (rx o lam)
        out[0,k]    += g2[0,k] - g2[1,k-1]
        out[Np-1,k] += g2[1,k] - g2[0,k+1]
    with the neighbour terms dropped at the domain ends for bc=inflow (there the boundary
    jump multiplies the prescribed inflow / is zeroed, AdvecRHS1D.m:15-16) and wrapped for
    bc=periodic."""
    Np = g.Np
    a_arr = np.asarray(a, dtype=float)
    ab = a_arr[..., None] if a_arr.ndim else float(a_arr)
    c0, c1 = face_coeff(ab, alpha)
    G = np.swapaxes(g.LIFT, 0, 1) @ lam                          # (..., 2, K)
    g0 = G[..., 0, :] * g.Fscale[0, :] * c0
    g1 = G[..., 1, :] * g.Fscale[1, :] * c1
    if bc == BC_INFLOW:
        g1 = g1.copy()
        g1[..., -1] = 0.0                                        # du(mapO) = 0
    out = -_bcast(a_arr) * (g.Dr.T @ (g.rx * lam))
    g1_left = np.roll(g1, 1, axis=-1)                            # g2[1,k-1]
    g0_right = np.roll(g0, -1, axis=-1)                          # g2[0,k+1]
    if bc != BC_PERIODIC:
        g1_left = g1_left.copy(); g0_right = g0_right.copy()
        g1_left[..., 0] = 0.0
        g0_right[..., -1] = 0.0
        if bc != BC_INFLOW:
            raise ValueError(bc)
    out[..., 0, :] += g0 - g1_left
    out[..., Np - 1, :] += g1 - g0_right
    return out


def adjoint_step(lam_u, lam_k, dt, a, g, alpha, bc, scheme=SCHEME_LSERK4):
    """Reverse of one low-storage RK step (stages s = last..0):
        lam_k += b_s lam_u ;  lam_u += dt L^T lam_k ;  lam_k *= a_s."""
    A, Bc, _ = scheme
    dt3 = _bcast(dt)
    for s in range(len(A) - 1, -1, -1):
        lam_k = lam_k + Bc[s] * lam_u
        lam_u = lam_u + dt3 * AdvecRHS1D_T(lam_k, a, g, alpha, bc)
        lam_k = A[s] * lam_k
    return lam_u, lam_k


FUNC_INT_U = 0    # J = int psi(x) u(x,T) dx, psi = 1 by default  (linear; cf. getK 'J=int(u)',
                  #                                          Main_finite_difference.py:153-155)
FUNC_INT_U2 = 1   # J = int u(x,T)^2 dx    (cf. 'J=int(u^2)', Main_finite_difference.py:225-227)


def quad_weights(g):
    """Nodal quadrature weights of the element mass matrix: (M_k 1)_i = J[i,k] * (Mref 1)_i."""
    Mref = ops.mass_matrix(g.V)
    return (Mref @ np.ones(g.Np))[:, None] * g.J                 # (Np, K)


def linear_weights(g, psi=None):
    """Weights jw[i,k] of a linear terminal functional J = sum jw o u  (psi: callable of x)."""
    w = quad_weights(g)
    return w if psi is None else w * psi(g.x)


def functional(u, g, kind, psi=None):
    if kind == FUNC_INT_U:
        return np.sum(linear_weights(g, psi) * u, axis=(-2, -1))
    if kind == FUNC_INT_U2:
        Mref = ops.mass_matrix(g.V)
        return np.sum(u * (g.J * (Mref @ u)), axis=(-2, -1))
    raise ValueError(kind)


def functional_grad(u, g, kind, psi=None):
    if kind == FUNC_INT_U:
        return np.broadcast_to(linear_weights(g, psi), u.shape).copy()
    if kind == FUNC_INT_U2:
        Mref = ops.mass_matrix(g.V)
        return 2.0 * g.J * (Mref @ u)
    raise ValueError(kind)


def adjoint_march(lamT, g, a, dt, nsteps, alpha=1.0, bc=BC_INFLOW, scheme=SCHEME_LSERK4, history=False):
    """Reverse-time march of the discrete adjoint of `advec_march` on operator set `g`.
    Returns (lam0, hist) with hist[n] = lam^n = dJ/du^n."""
    lam_u = np.array(lamT, dtype=float, copy=True)
    lam_k = np.zeros_like(lam_u)
    hist = [lam_u.copy()] if history else None
    for _ in range(nsteps):
        lam_u, lam_k = adjoint_step(lam_u, lam_k, dt, a, g, alpha, bc, scheme)
        if history:
            hist.append(lam_u.copy())
    if history:
        hist = np.stack(hist[::-1], axis=0)
    return lam_u, hist


def fwd_adj_indicator(u0, gc, gf, a, dt, nsteps, alpha=1.0, bc=BC_INFLOW, inflow=INFLOW_SIN_AT,
                      func=FUNC_INT_U, scheme=SCHEME_LSERK4, t0=0.0, psi=None):
    """The whole hot path for one (batch of) trajectory(ies), SURVEY App. E.5:

      1. coarse forward march on `gc` (order N)            -> u^n, n = 0..S
      2. terminal adjoint in the enriched space `gf` (order N+1): lam^S = dJ_f/du (P u^S)
      3. for n = S-1..0:  rho^n = P u^{n+1} - Phi_f(P u^n)        (fine one-step residual
                                                                    of the injected coarse solution)
                          eta[k] += sum_i lam^{n+1}[i,k] rho^n[i,k]
                          lam^n   = (dPhi_f)^T lam^{n+1}
    `lam` is the raw discrete adjoint dJ/du (mass included), so App. E.5's
    z^T M_k rho with z = M^-1 lam is the same number.  For a linear problem and linear J,
    sum_k eta[k] = J_f(P u^S) - J_f(u_f^S) exactly (tests check this effectivity).
    Returns dict(uT, J, lam0, eta, eta_scale) -- eta signed; consumers take |eta| (MAIN.m:51).
    eta is a sum of products lam*(P u^{n+1} - Phi_f(P u^n)) whose two parts cancel to
    O(h^{N+1}); any two fp64 evaluations agree to eps * eta_scale (the sum of the absolute
    values of the parts), not to eps * |eta| -- parity tests use eta_scale as the yardstick."""
    P = ops.prolongation(gc.N, gf.N)
    uT, hist = advec_march(u0, gc, a, dt, nsteps, alpha, bc, inflow, scheme, history=True, t0=t0)
    Jc = functional(uT, gc, func, psi)
    lam = functional_grad(P @ uT, gf, func, psi)
    lam_k = np.zeros_like(lam)
    eta = np.zeros(lam.shape[:-2] + (gc.K,))
    eta_scale = np.zeros_like(eta)   # sum of |terms| before cancellation (fp64 agreement scale)
    times = [t0 + 0.0 * np.asarray(dt, dtype=float)]
    for n in range(nsteps):
        times.append(times[-1] + dt)           # same accumulation as the forward march
    for n in range(nsteps - 1, -1, -1):
        t = times[n]
        uf, _ = step(P @ hist[n], np.zeros_like(lam), t, dt, a, gf, alpha, bc, inflow, scheme)
        rho = P @ hist[n + 1] - uf
        eta += np.sum(lam * rho, axis=-2)
        eta_scale += np.sum(np.abs(lam) * (np.abs(P @ hist[n + 1]) + np.abs(uf)), axis=-2)
        lam, lam_k = adjoint_step(lam, lam_k, dt, a, gf, alpha, bc, scheme)
    return dict(uT=uT, J=Jc, lam0=lam, eta=eta, hist=hist, eta_scale=eta_scale)


def rank_refine(eta, topk=1):
    """Refine flag / ranking: argmax of |eta| with lowest-index tie rule
    (np.argmax, Main_finite_difference.py:337; find(abs(err)==max(abs(err))), MAIN.m:137);
    ranking = stable descending sort (sort(...,'descend'), MAIN.m:99)."""
    ae = np.abs(eta)
    order = np.argsort(-ae, axis=-1, kind="stable")
    flags = np.zeros(ae.shape, dtype=np.uint8)
    np.put_along_axis(flags, order[..., :topk], 1, axis=-1)
    return order.astype(np.int32), flags
