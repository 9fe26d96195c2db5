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


def adapt_tdg(y0, tspan=(0.0, 2.0), Ks=2, n=1, iters=30, linear=False, device=0, B_global=None):
    """matlab/MAIN.m:19-166 (batched): Ks elements of order n, adjoint order n+1, refine the
    element with the largest batch-mean |err| by midpoint insertion (:137-141)."""
    s = TimeDG(linear=linear, device=device)
    times = np.linspace(tspan[0], tspan[1], Ks + 1)
    Ns = n * np.ones(Ks, dtype=int)
    hist = []
    for it in range(iters + 1):
        t1, y1, its = s.dg_march(Ns, Ks, times, y0)                # MAIN.m:32
        t2, v, err = s.adj_march(Ns + 1, Ks, times, y1, t1)        # MAIN.m:34
        mean_err = _batch_reduce(s, err, B_global)                 # mean_b |err| (MAIN.m:51 abs)
        times_new, Ns_new, ref_i = tdg_refine(times, Ns, mean_err, n)
        hist.append(dict(it=it, times=times.copy(), err=mean_err, ref_idx=ref_i, err_total=float(mean_err.sum()),
                         max_newton_its=int(its.max()), yT_mean=float(y1[:, -1, -1].mean())))
        times, Ns, Ks = times_new, Ns_new, Ks + 1
    s.close()
    return hist
