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


LAST_LOOP_DEVICE_MS = {}    # CUDA-event time of the last device-resident loop (enqueue of the first kernel .. last kernel done)


def _distributed():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def _fd_device_loop(s, u0, times, iters):
    """dgadj_fd_adapt_loop: every iteration enqueued by one call, the mesh lives on the device."""
    torch = s.torch
    u0 = u0.contiguous().view(-1)
    B, n0 = u0.numel(), times.size - 1
    nmax, W = n0 + iters, n0 + iters + 2
    times = np.ascontiguousarray(times, dtype=np.float64)
    th = torch.zeros((iters + 1, W), dtype=torch.float64, device=u0.device)
    eh = torch.zeros((iters + 1, nmax), dtype=torch.float64, device=u0.device)
    ri = torch.zeros(iters + 1, dtype=torch.int32, device=u0.device)
    tot = torch.zeros(iters + 1, dtype=torch.float64, device=u0.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = s.lib.dgadj_fd_adapt_loop(s._h, B, iters, n0, s.ref_factor, _lib.FD_ODE[s.ode], _lib.FD_FUNCTIONAL[s.functional],
                                   C.c_void_p(times.ctypes.data), C.c_void_p(u0.data_ptr()), C.c_void_p(th.data_ptr()),
                                   C.c_void_p(eh.data_ptr()), C.c_void_p(ri.data_ptr()), C.c_void_p(tot.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream(s.device).cuda_stream))
    e1.record()
    if rc != _lib.OK:
        raise _lib.DgadjError(rc, s.lib.dgadj_last_error(s._h).decode())
    e1.synchronize()
    LAST_LOOP_DEVICE_MS["fd"] = e0.elapsed_time(e1)
    th, eh, ri, tot = th.cpu().numpy(), eh.cpu().numpy(), ri.cpu().numpy(), tot.cpu().numpy()   # the one read-back
    return [dict(it=it, times=th[it, :n0 + it + 1].copy(), err_steps=eh[it, :n0 + it].copy(), ref_idx=int(ri[it]),
                 err_total=float(tot[it])) for it in range(iters + 1)]


def adapt_fd_per_trajectory(u0, tspan=(0.0, 2.0), n_steps=2, iters=30, ode="sin", functional="int_u2", ref_factor=4, device=0):
    """python/Main_finite_difference.py:263-343 for every trajectory of a batch at once, each on ITS OWN mesh (the
    reference's single-trajectory loop, no batch rule): `dgadj_fd_adapt_loop_pt`, one kernel, each thread runs the whole
    loop.  Returns dict(times[B, n_steps+iters+1], ref_idx[B, iters+1], err_total[B, iters+1]) as NumPy arrays."""
    s = FDAdjoint(ode=ode, functional=functional, ref_factor=ref_factor, device=device)
    torch = s.torch
    u0 = u0.contiguous().view(-1)
    B = u0.numel()
    times0 = np.ascontiguousarray(np.linspace(tspan[0], tspan[1], n_steps + 1))
    to = torch.zeros((B, n_steps + iters + 1), dtype=torch.float64, device=u0.device)
    ri = torch.zeros((B, iters + 1), dtype=torch.int32, device=u0.device)
    tot = torch.zeros((B, iters + 1), dtype=torch.float64, device=u0.device)
    rc = s.lib.dgadj_fd_adapt_loop_pt(s._h, B, iters, n_steps, s.ref_factor, _lib.FD_ODE[s.ode], _lib.FD_FUNCTIONAL[s.functional],
                                      C.c_void_p(times0.ctypes.data), C.c_void_p(u0.data_ptr()), C.c_void_p(to.data_ptr()),
                                      C.c_void_p(ri.data_ptr()), C.c_void_p(tot.data_ptr()),
                                      C.c_void_p(torch.cuda.current_stream(s.device).cuda_stream))
    if rc != _lib.OK:
        raise _lib.DgadjError(rc, s.lib.dgadj_last_error(s._h).decode())
    out = dict(times=to.cpu().numpy(), ref_idx=ri.cpu().numpy(), err_total=tot.cpu().numpy())
    s.close()
    return out


def adapt_tdg_per_trajectory(y0, tspan=(0.0, 2.0), Ks=2, n=1, iters=30, linear=False, device=0, quirks=True):
    """matlab/MAIN.m:29-166 for every trajectory of a batch at once, each on ITS OWN mesh (`dgadj_tdg_adapt_loop_pt`):
    one warp per trajectory builds its element blocks T0 + h T1, marches, solves the adjoint at order n+1, refines its own
    argmax element.  Returns dict(times[B, Ks+iters+1], ref_idx[B, iters+1], err_total[B, iters+1], y_last[B, Ks+iters, n+1],
    its_last[B, Ks+iters]) -- the first three as NumPy arrays."""
    s = TimeDG(linear=linear, device=device, quirks=quirks)
    torch = s.torch
    y0 = y0.contiguous().view(-1)
    B, Kmax, W, Np = y0.numel(), Ks + iters, Ks + iters + 2, n + 1
    tp = _tdg_templates(s, n)
    times0 = np.linspace(tspan[0], tspan[1], Ks + 1)
    p = lambda a: C.c_void_p(a.ctypes.data)
    args = _lib.TdgLoopArgs(B=B, iters=iters, Ks0=Ks, Np=Np, nq_march=tp["nq_m"], nq_adj=tp["nq_a"], linear=int(s.linear),
                            maxit=s.maxit, y0_per_trajectory=1, tol=s.tol, y0_hard=1.0, times0_host=None,
                            march_T0_host=p(tp["m0"]), march_T1_host=p(tp["m1"]), adj_T0_host=p(tp["a0"]), adj_T1_host=p(tp["a1"]))
    kw = dict(dtype=torch.float64, device=y0.device)
    tm = torch.zeros((B, W), **kw)
    tm[:, :Ks + 1] = torch.as_tensor(times0, device=y0.device)
    ri = torch.zeros((B, iters + 1), dtype=torch.int32, device=y0.device)
    tot = torch.zeros((B, iters + 1), **kw)
    yl = torch.zeros((B, Kmax, Np), **kw)
    il = torch.zeros((B, Kmax), dtype=torch.int32, device=y0.device)
    d = lambda t: C.c_void_p(t.data_ptr())
    rc = s.lib.dgadj_tdg_adapt_loop_pt(s._h, C.byref(args), d(y0), d(tm), d(ri), d(tot), d(yl), d(il), s._stream())
    if rc != _lib.OK:
        raise _lib.DgadjError(rc, s.lib.dgadj_last_error(s._h).decode())
    out = dict(times=tm[:, :Kmax + 1].cpu().numpy(), ref_idx=ri.cpu().numpy(), err_total=tot.cpu().numpy(), y_last=yl, its_last=il)
    s.close()
    return out


def adapt_fd(u0, tspan=(0.0, 2.0), n_steps=2, iters=30, ode="sin", functional="int_u2", ref_factor=4,
             tol=None, device=0, B_global=None, device_loop=None):
    """python/Main_finite_difference.py:263-343 (batched): start from n_steps uniform steps,
    refine the step with the largest batch-mean indicator `iters` times (or until the summed
    indicator drops below tol, :263).  Returns the history: one dict per iteration with
    times, err_steps (batch mean), ref_idx (0-based element refined), err_total.
    device_loop (default: whenever there is no tolerance to test and no other rank to meet): the whole loop
    is one C-ABI call (`dgadj_fd_adapt_loop`) -- mesh, tables, argmax and midpoint insertion on the device,
    one read-back at the end; otherwise the host drives it iteration by iteration."""
    s = FDAdjoint(ode=ode, functional=functional, ref_factor=ref_factor, device=device)
    times = np.linspace(tspan[0], tspan[1], n_steps + 1)
    if device_loop is None:
        device_loop = tol is None and B_global is None and not _distributed()
    if device_loop:
        hist = _fd_device_loop(s, u0, times, iters)
        s.close()
        return hist
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


_TEMPLATES = {}     # (order, linear, quirks) -> template blocks: host work done once per process


def _tdg_templates(s, n):
    key = (int(n), s.linear, s.quirks)
    if key not in _TEMPLATES:
        _TEMPLATES[key] = _build_tdg_templates(s, n)
    return _TEMPLATES[key]


def _build_tdg_templates(s, n):
    """Element blocks of the time-DG march / adjoint as affine functions of the element width,
    block(h) = T0 + h T1: the host builds the blocks of the second element of the meshes [0, 1, 2] and
    [0, 2, 4] (fem_setup operators, the polyfit/polyval matrices, the mirrored quadrature interval: everything
    `TimeDG.march_constants` / `adjoint_constants` produce) and differences them."""
    out = {}
    for kind in ("m", "a"):
        blocks = []
        for hh in (1.0, 2.0):
            times = np.array([0.0, hh, 2 * hh])
            Ns = n * np.ones(2, dtype=int)
            cm, nodes, NP, nqm = s.march_constants(Ns, times)
            if kind == "m":
                blocks.append(cm.reshape(2, -1)[1].copy())
                out["nq_m"] = nqm
            else:
                ca, _, _, nqa = s.adjoint_constants(Ns + 1, nodes)
                blocks.append(ca.reshape(2, -1)[1].copy())
                out["nq_a"] = nqa
        out[kind + "1"] = np.ascontiguousarray(blocks[1] - blocks[0])
        out[kind + "0"] = np.ascontiguousarray(2.0 * blocks[0] - blocks[1])
    return out


def _tdg_device_loop(s, y0, times, n, iters):
    """dgadj_tdg_adapt_loop: every iteration enqueued by one call, the mesh lives on the device."""
    torch = s.torch
    y0 = y0.contiguous().view(-1)
    B, Ks0 = y0.numel(), times.size - 1
    Kmax, W, Np = Ks0 + iters, Ks0 + iters + 2, n + 1
    tp = _tdg_templates(s, n)
    times = np.ascontiguousarray(times, dtype=np.float64)
    p = lambda a: C.c_void_p(a.ctypes.data)
    args = _lib.TdgLoopArgs(B=B, iters=iters, Ks0=Ks0, Np=Np, nq_march=tp["nq_m"], nq_adj=tp["nq_a"], linear=int(s.linear),
                            maxit=s.maxit, y0_per_trajectory=1, tol=s.tol, y0_hard=1.0, times0_host=p(times),
                            march_T0_host=p(tp["m0"]), march_T1_host=p(tp["m1"]), adj_T0_host=p(tp["a0"]), adj_T1_host=p(tp["a1"]))
    kw = dict(dtype=torch.float64, device=y0.device)
    th, eh = torch.zeros((iters + 1, W), **kw), torch.zeros((iters + 1, Kmax), **kw)
    ri = torch.zeros(iters + 1, dtype=torch.int32, device=y0.device)
    st, ist = torch.zeros((iters + 1, 2), **kw), torch.zeros((iters + 1, 3), dtype=torch.int32, device=y0.device)
    d = lambda t: C.c_void_p(t.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = s.lib.dgadj_tdg_adapt_loop(s._h, C.byref(args), d(y0), d(th), d(eh), d(ri), d(st), d(ist), C.c_void_p(0), s._stream())
    e1.record()
    if rc != _lib.OK:
        raise _lib.DgadjError(rc, s.lib.dgadj_last_error(s._h).decode())
    e1.synchronize()
    LAST_LOOP_DEVICE_MS["tdg"] = e0.elapsed_time(e1)
    th, eh, ri, st, ist = (t.cpu().numpy() for t in (th, eh, ri, st, ist))        # the one read-back
    return [dict(it=it, times=th[it, :Ks0 + it + 1].copy(), err=eh[it, :Ks0 + it].copy(), ref_idx=int(ri[it]),
                 err_total=float(st[it, 1]), max_newton_its=int(ist[it, 0]), yT_mean=float(st[it, 0]),
                 not_converged=int(ist[it, 1]), non_finite=int(ist[it, 2])) for it in range(iters + 1)]


def adapt_tdg(y0, tspan=(0.0, 2.0), Ks=2, n=1, iters=30, linear=False, device=0, B_global=None, quirks=True,
              device_loop=None):
    """matlab/MAIN.m:19-166 (batched): Ks elements of order n, adjoint order n+1, refine the
    element with the largest batch-mean |err| by midpoint insertion (:137-141).  Every trajectory's
    first-element residual is measured against its own initial value (the reference runs one
    trajectory with y0 = 1 hard-coded in adj_march.m:9 -- identical for that case).  quirks=False:
    see TimeDG.
    device_loop (default: whenever no other rank has to be met): the whole loop is one C-ABI call
    (`dgadj_tdg_adapt_loop`): element blocks, march, adjoint, batch mean, argmax and midpoint insertion on
    the device, one read-back at the end; otherwise the host drives it iteration by iteration."""
    s = TimeDG(linear=linear, device=device, quirks=quirks)
    times = np.linspace(tspan[0], tspan[1], Ks + 1)
    if device_loop is None:
        device_loop = B_global is None and not _distributed()
    if device_loop:
        hist = _tdg_device_loop(s, y0, times, n, iters)
        s.close()
        return hist
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
                device=0, B_global=None, ic_term=True, keep_indicators=True, orders=None):
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
    initial data.  orders: per-element orders N_k <= N of the initial mesh (hp; `Ns(k)`, matlab/MAIN.m:21) --
    the two halves of a split element keep its order (MAIN.m:141 gives the new element the order of the run).
    Returns the history (mesh, mean indicator, refined elements, J mean, the signed batch-mean estimate, the
    element orders) per iteration."""
    import torch
    v_x = np.asarray(v_x, dtype=np.float64)
    hist = []
    # ONE handle for the whole loop (capacity = the last mesh); the vertex list also lives on the device, where the
    # refinement rule runs (`dgadj_refine_shared`: argmax / top-k + midpoint insertion).  The march length of the
    # next iteration follows the CFL rule of the refined mesh, so the host mirrors the mesh from the indices of the
    # split elements -- 4 topk bytes, the one read-back per iteration.
    s = AdvecDG1D(N, v_x=v_x, alpha=alpha, bc=bc, inflow=inflow, psi=psi, device=device, capacity=v_x.size - 1 + iters * topk)
    dev = torch.device("cuda", device)
    d_vx = torch.zeros(s.capacity + topk + 1, dtype=torch.float64, device=dev)
    d_vx[:v_x.size] = torch.as_tensor(v_x, device=dev)
    orders = None if orders is None else np.asarray(orders, dtype=np.int64).copy()
    if orders is not None and orders.size != v_x.size - 1:
        raise ValueError("orders: one entry per element of the initial mesh")
    for it in range(iters + 1):
        if it:
            s.set_mesh(v_x)
        if orders is not None:
            s.set_element_orders(orders)
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
        refined = s.refine_shared(sums[:K], d_vx, topk)                # device: argmax / top-k, midpoint insertion
        order = np.sort(refined.cpu().numpy())                         # the read-back: which elements were split
        mean_eta = (sums[:K] / float(Bg)).cpu().numpy() if keep_indicators else None
        hist.append(dict(it=it, v_x=v_x.copy(), K=K, S=S, mean_eta=mean_eta, refined=order,
                         eta_total=float(sums[K]) / float(Bg) if keep_indicators else None,
                         J_mean=float(sums[K + 3]) / float(Bg) if keep_indicators else None,
                         J=out["J"], estimate=eta.sum(1), orders=None if orders is None else orders.copy()))
        mids = 0.5 * (v_x[order] + v_x[order + 1])                     # host mirror of the device mesh (same arithmetic)
        v_x = np.sort(np.concatenate([v_x, mids]))
        if orders is not None:
            orders = np.insert(orders, order + 1, orders[order])       # both halves keep the parent's order
    if not np.array_equal(d_vx[:v_x.size].cpu().numpy(), v_x):       # the device mesh and its host mirror
        raise RuntimeError("adapt_advec: the host mirror of the mesh has left the device mesh")
    s.close()
    return hist
