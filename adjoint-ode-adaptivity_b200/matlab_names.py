"""The reference's MATLAB entry points under their own names AND argument lists:

    rhsu        = AdvecRHS1D(u, timelocal, a)                     utils/AdvecRHS1D.m:1
    ulimit      = SlopeLimitN(u)                                  utils/SlopeLimitN.m:1
    ulimit      = SlopeLimit1(u)                                  utils/SlopeLimit1.m:1
    mfunc       = minmod(v)                                       utils/minmod.m:1   (v is [m, K], reduces rows)
    [t, y]      = dg_march(Ns, Ks, times, y0, x_true, u_true)     matlab/dg_march.m:1
    [t, v, err] = adj_march(Ns, Ks, times)                        matlab/adj_march.m:1
    [t, v, err] = adj_rec(Ns, Ks, times)                          matlab/adj_rec.m:1
    [err, res]  = err_contribution(Ks, Ns, uh, t1)                matlab/err_contribution.m:1
    [t, y]      = fwd_euler_march(y0, times, ode_fn)              matlab/fwd_euler_march.m:1 (a broken stub in
                                                                  the reference; semantics of forwardSolve,
                                                                  python/Main_finite_difference.py:34-51)

The MATLAB routines share their state through globals (`Globals1D`, utils/Globals1D.m:3-17; the primal
`y1 t1` that adj_march reads, matlab/adj_march.m:4).  The analogue here is the module-level state `G`: the
device handles, created by `StartUp1D(...)` (for the PDE routines) or on first use (the ODE routines), and the
primal of the last `dg_march`.  Arrays are float64 CUDA tensors (or NumPy arrays, copied); a leading batch axis
is allowed.  The older handle-first forms (`AdvecRHS1D(solver, u, t, a)` ...) keep working.
"""
from __future__ import annotations

import numpy as np

from . import fd as _fd
from .burgers import BurgersDG1D
from .solver import AdvecDG1D
from .tdg import TimeDG


class _Globals:
    """utils/Globals1D.m:3-17 (the PDE handles) + the globals of matlab/MAIN.m (y1, t1: the last primal)."""
    advec = None
    burgers = None
    tdg = None
    y1 = t1 = y0 = its = None
    per_trajectory_y0 = False     # False: adj_march measures the first element against y0 = 1 (adj_march.m:9)


G = _Globals()


def Globals1D():
    return G


def StartUp1D(N, K=None, domain=(0.0, 1.0), v_x=None, alpha=1.0, bc="inflow", inflow="sin_at", device=0):
    """utils/StartUp1D.m:5-39 for the device: builds the operators of the mesh and the handles the PDE
    routines below use (AdvecRHS1D; SlopeLimitN / SlopeLimit1 with ghost averages copied at the ends, quirk C-16,
    unless bc = "periodic")."""
    for h in (G.advec, G.burgers):
        if h is not None:
            h.close()
    G.advec = AdvecDG1D(N, K, domain=domain, v_x=v_x, alpha=alpha, bc=bc, inflow=inflow, device=device)
    G.burgers = BurgersDG1D(N, K, domain=domain, v_x=v_x, bc="periodic" if bc == "periodic" else "free", device=device)
    return G


def set_time_dg(linear=False, quirks=True, device=0, per_trajectory_y0=False, tol=1e-7, maxit=500):
    """The switches the reference sets by editing its files: `linear` (dg_march.m:8-9, adj_march.m:12), the
    Newton tolerance / iteration cap (dg_march.m:36,44)."""
    if G.tdg is not None:
        G.tdg.close()
    G.tdg = TimeDG(linear=linear, device=device, tol=tol, maxit=maxit, quirks=quirks)
    G.per_trajectory_y0 = bool(per_trajectory_y0)
    return G


def _is_handle(x):
    return isinstance(x, (AdvecDG1D, BurgersDG1D, TimeDG))


def _need(h, what):
    if h is None:
        raise RuntimeError(f"{what}: call StartUp1D(N, K, ...) first (the MATLAB routines read Globals1D)")
    return h


def _dev(u, device=0):
    import torch
    if isinstance(u, torch.Tensor):
        return u, True
    return torch.as_tensor(np.asarray(u, dtype=np.float64), device=torch.device("cuda", device)), False


def AdvecRHS1D(*args):
    """rhsu = AdvecRHS1D(u, timelocal, a)"""
    if _is_handle(args[0]):
        solver, u, timelocal, a = args
        return solver.rhs(u, timelocal, a)
    u, timelocal, a = args
    s = _need(G.advec, "AdvecRHS1D")
    d_u, was = _dev(u, s.device)
    squeeze = d_u.ndim == 2
    out = s.rhs(d_u, timelocal, a)
    out = out[0] if squeeze else out
    return out if was else out.cpu().numpy()


def _limit(kind, *args):
    if _is_handle(args[0]):
        return args[0].slope_limit(args[1], kind=kind)
    s = _need(G.burgers, "SlopeLimit" + kind)
    d_u, was = _dev(args[0], s.device)
    squeeze = d_u.ndim == 2
    out = s.slope_limit(d_u, kind=kind)
    out = out[0] if squeeze else out
    return out if was else out.cpu().numpy()


def SlopeLimitN(*args):
    """ulimit = SlopeLimitN(u)"""
    return _limit("N", *args)


def SlopeLimit1(*args):
    """ulimit = SlopeLimit1(u)"""
    return _limit("1", *args)


def minmod(v):
    """mfunc = minmod(v)  (utils/minmod.m:6-12: v is [m, K]; s = sum(sign(v))/m; |s| == 1 -> s min|v|, else 0).
    Host helper: inside the marches the three-argument form is fused into the limiter kernels."""
    v = np.asarray(v, dtype=np.float64)
    m = v.shape[0]
    s = np.sum(np.sign(v), axis=0) / m
    out = np.zeros(v.shape[1:])
    ids = np.abs(s) == 1
    out[ids] = (s * np.min(np.abs(v), axis=0))[ids]
    return out


def dg_march(*args):
    """[t, y] = dg_march(Ns, Ks, times, y0, x_true, u_true).  y0: a scalar (the reference) or [B] values.
    The primal is also left in the module state (`G.y1`, `G.t1`, as the globals of matlab/MAIN.m:32), with the
    Newton counts `G.its` and the per-trajectory status word `G.status` (bit 0: a solve did not converge --
    the reference prints that, dg_march.m:69-73 --, bit 1: non-finite values)."""
    import torch
    if _is_handle(args[0]):
        tdg, args = args[0], args[1:]
    else:
        if G.tdg is None:
            set_time_dg()
        tdg = G.tdg
    Ns, Ks, times, y0 = args[:4]
    d_y0, was = _dev(np.atleast_1d(y0) if not isinstance(y0, torch.Tensor) else y0, tdg.device)
    t, y, its = tdg.dg_march(Ns, Ks, times, d_y0.view(-1))
    G.t1, G.y1, G.y0, G.its, G.status = t, y, d_y0.view(-1), its, tdg.status()
    G._tdg_used = tdg
    return t, y


def adj_march(*args):
    """[t, v, err] = adj_march(Ns, Ks, times): the primal is the one the last dg_march left behind
    (matlab/adj_march.m:4 reads the globals y1, t1); Ns = the adjoint orders (matlab/MAIN.m:34 passes Ns+1)."""
    if _is_handle(args[0]):
        tdg, Ns, Ks, times, y1, t1 = args
        return tdg.adj_march(Ns, Ks, times, y1, t1)
    Ns, Ks, times = args[:3]
    if G.y1 is None:
        raise RuntimeError("adj_march: no primal -- call dg_march first (adj_march.m:4 reads the globals y1, t1)")
    return G._tdg_used.adj_march(Ns, Ks, times, G.y1, G.t1, y0=G.y0 if G.per_trajectory_y0 else 1.0)


def adj_rec(*args):
    """[t, v, err] = adj_rec(Ns, Ks, times)  (takes the PRIMAL orders, matlab/MAIN.m:35)"""
    if _is_handle(args[0]):
        tdg, Ns, Ks, times, y1, t1 = args
        return tdg.adj_rec(Ns, Ks, times, y1, t1)
    Ns, Ks, times = args[:3]
    if G.y1 is None:
        raise RuntimeError("adj_rec: no primal -- call dg_march first")
    return G._tdg_used.adj_rec(Ns, Ks, times, G.y1, G.t1, y0=G.y0 if G.per_trajectory_y0 else 1.0)


def err_contribution(*args):
    """[err, res] = err_contribution(Ks, Ns, uh, t1)"""
    if _is_handle(args[0]):
        tdg, Ks, Ns, uh, t1 = args
    else:
        Ks, Ns, uh, t1 = args
        if G.tdg is None:
            set_time_dg()
        tdg = G.tdg
    return tdg.err_contribution(Ks, Ns, uh, t1), [None] * Ks


def fwd_euler_march(y0, times, ode_fn="sin", device=0):
    """[t, y] = fwd_euler_march(y0, times, ode_fn): explicit Euler on the mesh `times`.  ode_fn: "sin" / "linear",
    or a callable f(u) that is PROBED against the reference's two right-hand sides (sin(u), u)."""
    if callable(ode_fn):
        r = float(ode_fn(0.3))
        if np.isclose(r, np.sin(0.3), rtol=1e-14, atol=0):
            ode_fn = "sin"
        elif np.isclose(r, 0.3, rtol=1e-14, atol=0):
            ode_fn = "linear"
        else:
            raise NotImplementedError("ode_fn is neither sin(u) nor u: the device path cannot run an arbitrary callable")
    s = _fd._default_handle(ode_fn, "int_u2", 4, device)
    import torch
    d_y0, was = _dev(np.atleast_1d(y0) if not isinstance(y0, torch.Tensor) else y0, device)
    y = s.forwardSolve(np.diff(np.asarray(times, dtype=np.float64)), d_y0.view(-1))
    return np.asarray(times, dtype=np.float64), (y if was else y.cpu().numpy())
