"""Adjoint-driven adaptive refinement loops (BASELINE config 5) on top of the device kernels:
the DG-in-time loop of matlab/MAIN.m:29-166 and the finite-difference loop of
python/Main_finite_difference.py:263-343, batched over initial conditions with the shared-mesh /
batch-mean-indicator rule of python/Main_variable_params.py:330-344 (for a batch of one this
is exactly the reference's single-trajectory loop).

Per iteration: march -> adjoint -> per-element indicator (device), fixed-order batch reduction
(device, `dgadj_reduce_indicators`; all-reduced across ranks when torch.distributed is
initialised), argmax + midpoint insertion (host: a K-long vector).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .fd import FDAdjoint, refine_mesh
from .sharding import allreduce_indicators
from .tdg import TimeDG, refine as tdg_refine
from .solver import AdvecDG1D


def _batch_reduce(obj, eta, B_global=None):
    """mean_b |eta[b, k]| through the fixed-order device reduction (+ ordered all-reduce)."""
    torch = obj.torch
    eta = eta.contiguous()
    B, K = eta.shape
    sums = torch.empty(K + 4, dtype=torch.float64, device=eta.device)
    rc = obj.lib.dgadj_reduce_indicators(obj._h, B, K, C.c_void_p(eta.data_ptr()), C.c_void_p(0),
                                         C.c_void_p(sums.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream(obj.device).cuda_stream))
    if rc != _lib.OK:
        raise _lib.DgadjError(rc, obj.lib.dgadj_last_error(obj._h).decode())
    sums = allreduce_indicators(sums, ordered=True)
    Bg = B_global if B_global is not None else B
    return (sums[:K] / float(Bg)).cpu().numpy()


def adapt_fd(u0, tspan=(0.0, 2.0), n_steps=2, iters=30, ode="sin", functional="int_u2", ref_factor=4,
             tol=None, device=0, B_global=None):
    """python/Main_finite_difference.py:263-343 (batched): start from n_steps uniform steps,
    refine the step with the largest batch-mean indicator `iters` times (or until the summed
    indicator drops below tol, :263).  Returns the history: one dict per iteration with
    times, err_steps (batch mean), ref_idx (0-based element refined), err_total."""
    s = FDAdjoint(ode=ode, functional=functional, ref_factor=ref_factor, device=device)
    times = np.linspace(tspan[0], tspan[1], n_steps + 1)
    hist = []
    for it in range(iters + 1):
        out = s.solve(u0, np.diff(times), want=("err_steps",))
        mean_steps = _batch_reduce(s, out["err_steps"], B_global)
        ref_idx = int(np.argmax(mean_steps))                       # :337 (np.argmax, 0-based element)
        hist.append(dict(it=it, times=times.copy(), err_steps=mean_steps, ref_idx=ref_idx,
                         err_total=float(mean_steps.sum())))
        if tol is not None and hist[-1]["err_total"] <= tol:
            break
        times = refine_mesh(times, ref_idx)                        # :336-341
    s.close()
    return hist


def adapt_tdg(y0, tspan=(0.0, 2.0), Ks=2, n=1, iters=30, linear=False, device=0, B_global=None, quirks=True):
    """matlab/MAIN.m:19-166 (batched): Ks elements of order n, adjoint order n+1, refine the
    element with the largest batch-mean |err| by midpoint insertion (:137-141).  Every trajectory's
    first-element residual is measured against its own initial value (the reference runs one
    trajectory with y0 = 1 hard-coded in adj_march.m:9 -- identical for that case).  quirks=False:
    see TimeDG."""
    s = TimeDG(linear=linear, device=device, quirks=quirks)
    times = np.linspace(tspan[0], tspan[1], Ks + 1)
    Ns = n * np.ones(Ks, dtype=int)
    hist = []
    for it in range(iters + 1):
        t1, y1, its = s.dg_march(Ns, Ks, times, y0)                # MAIN.m:32
        t2, v, err = s.adj_march(Ns + 1, Ks, times, y1, t1, y0=y0)  # MAIN.m:34
        mean_err = _batch_reduce(s, err, B_global)                 # mean_b |err| (MAIN.m:51 abs)
        times_new, Ns_new, ref_i = tdg_refine(times, Ns, mean_err, n)
        hist.append(dict(it=it, times=times.copy(), err=mean_err, ref_idx=ref_i, err_total=float(mean_err.sum()),
                         max_newton_its=int(its.max()), yT_mean=float(y1[:, -1, -1].mean())))
        times, Ns, Ks = times_new, Ns_new, Ks + 1
    s.close()
    return hist


def adapt_advec(u0_fn, N, v_x, a, T, iters=10, topk=1, alpha=0.0, bc="periodic", inflow="zero", cfl=0.25, psi=None,
                device=0, B_global=None, ic_term=True):
    """Adjoint-driven h-refinement of the DG-in-space advection march (BASELINE config 5 for the
    PDE path: non-uniform h): per iteration the batch is marched forward and backward on the
    current mesh, the per-element indicators are reduced over the batch in a fixed order, and the
    `topk` elements with the largest batch-mean |eta| are split at their midpoints
    (matlab/MAIN.m:137-141 applied to space; batch rule python/Main_variable_params.py:340-341).

    u0_fn(x) -> float64 CUDA tensor [B, Np, K] of initial conditions at the nodes x[Np, K]
    (re-evaluated on every mesh).  ic_term: the march's indicator weighs the residual of the
    *evolution* (coarse and enriched solutions start from the same interpolated data); with
    ic_term the interpolation defect of the initial data enters too, weighted by the adjoint at
    t = 0:  eta_k += lam0_k . (P u0(x_c) - u0(x_f))_k  -- so that sum_k eta_k is the whole difference
    between the functional of the coarse solution and that of the order-N+1 solution of the true
    initial data.  Returns the history (mesh, mean indicator, refined elements, J mean, the signed
    batch-mean estimate) per iteration."""
    import torch
    v_x = np.asarray(v_x, dtype=np.float64)
    hist = []
    for it in range(iters + 1):
        s = AdvecDG1D(N, v_x=v_x, alpha=alpha, bc=bc, inflow=inflow, psi=psi, device=device)
        xmin = np.min(np.abs(s.g.x[0, :] - s.g.x[1, :]))
        S = int(np.ceil(T / (cfl * xmin / abs(a))))
        dt = T / S
        u0 = u0_fn(s.g.x)
        out = s.fwd_adj(u0, a, dt, S, want_uT=False, want_lam0=ic_term, window="auto")   # windows only if the ring cannot fit
        eta = out["eta"]
        if ic_term:
            s.ic_indicator(u0, u0_fn(s.gf.x), out["lam0"], eta)        # dgadj_ic_indicator, in place
        sums = allreduce_indicators(s.reduce_indicators(eta, out["J"]), ordered=True)
        Bg = B_global if B_global is not None else u0.shape[0]
        K = s.K
        mean_eta = (sums[:K] / float(Bg)).cpu().numpy()
        order = np.argsort(-mean_eta, kind="stable")[:topk]            # ties by lowest index
        hist.append(dict(it=it, v_x=v_x.copy(), K=K, S=S, mean_eta=mean_eta, refined=np.sort(order),
                         eta_total=float(mean_eta.sum()), J_mean=float(sums[K + 3]) / float(Bg),
                         J=out["J"], estimate=eta.sum(1)))
        s.close()
        mids = 0.5 * (v_x[order] + v_x[order + 1])
        v_x = np.sort(np.concatenate([v_x, mids]))
    return hist
