"""Python host of the dgadj C-ABI: batched 1-D advection DG forward march, discrete adjoint
march, per-element adjoint-weighted error indicator and refine ranking on a B200.

Array layouts are the reference's: fields are (B, Np, K) C-order float64 -- the batched form
of galerkin.py's (Np, K) arrays (python/galerkin.py:216).  NumPy arrays go through the
`*_host` entry points (chunked H2D / kernels / D2H inside the library); torch CUDA tensors
(or any `__dlpack__` CUDA object) go through the device entry points on torch's current
stream.  torch is used for device memory only.

Reference routines behind each method: see include/dgadj.h.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np

from . import _lib
from .galerkin import BaseGalerkin1D

TWO_PI = 2.0 * math.pi


def _torch():
    import torch
    return torch


def _is_host(x):
    return isinstance(x, np.ndarray) or np.isscalar(x) or isinstance(x, (list, tuple))


def _as_device_tensor(x):
    """torch CUDA tensor view of x (torch tensor or DLPack exporter), fp64, contiguous."""
    torch = _torch()
    if not isinstance(x, torch.Tensor):
        x = torch.from_dlpack(x)
    if not x.is_cuda:
        raise TypeError("device path needs a CUDA tensor; pass a numpy array for the host path")
    if x.dtype != torch.float64:
        raise TypeError(f"dgadj computes in float64, got {x.dtype}")
    return x.contiguous()


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _np_ptr(a):
    return C.c_void_p(0) if a is None else C.c_void_p(a.ctypes.data)


class AdvecDG1D:
    """One mesh + operator set + device handle.

    Parameters mirror the reference's knobs: `alpha` / `a` of AdvecRHS1D (utils/AdvecRHS1D.m:1,9),
    inflow boundary data (`sin_at` = AdvecRHS1D.m:14, `sin_aat` = utils/One_code.mlx), the mesh of
    MeshGen1D (domain, K) or an arbitrary vertex list `v_x` (refined meshes, matlab/MAIN.m:138-141).
    The adjoint is solved one order higher (matlab/MAIN.m:34).
    """

    def __init__(self, N: int, K: Optional[int] = None, domain=(0.0, 1.0), v_x=None, alpha: float = 1.0,
                 bc: str = "inflow", inflow: str = "sin_at", functional: str = "int_u",
                 scheme: str = "lserk4", device: int = 0, psi=None, capacity: Optional[int] = None):
        """capacity: the largest mesh (elements) this handle will ever hold (default: the initial mesh); a
        refinement loop creates the handle once and moves it from mesh to mesh with `set_mesh`."""
        self.lib = _lib.load()
        g0 = BaseGalerkin1D(n=N, k=K, domain=domain, v_x=v_x)
        self.N, self.Np, self.NpF = N, N + 1, N + 2
        self.alpha, self.bc, self.inflow, self.functional, self.scheme = alpha, bc, inflow, functional, scheme
        self.device = device
        self.capacity = max(int(capacity or 0), g0.k)
        cfg = _lib.Config(device=device, N=N, K=self.capacity, bc=_lib.BC[bc], inflow=_lib.INFLOW[inflow],
                          functional=_lib.FUNCTIONAL[functional], scheme=_lib.SCHEME[scheme], reserved=0,
                          alpha=float(alpha))
        self._h = C.c_void_p(0)
        rc = self.lib.dgadj_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.OK:
            self._h = C.c_void_p(0)
            raise _lib.DgadjError(rc, "dgadj_create failed (an sm_100 device is required; there is no CPU path)")
        self._psi = psi
        self.set_mesh(g=g0)

    def set_mesh(self, v_x=None, g=None):
        """Move the handle to the mesh with vertices v_x (at most `capacity` elements): the operators of both
        spaces (StartUp1D chain, host) and the functional weights are set again on the SAME device handle
        (matlab/MAIN.m:138-141 applied to space: the refined mesh of the next iteration)."""
        N = self.N
        self.g = g if g is not None else BaseGalerkin1D(n=N, v_x=np.asarray(v_x, dtype=np.float64))
        self.gf = BaseGalerkin1D(n=N + 1, v_x=self.g.v_x)
        self.K = self.g.k
        if self.K > self.capacity:
            raise ValueError(f"mesh of {self.K} elements exceeds the handle's capacity {self.capacity}")
        g, gf = self.g, self.gf
        psi = self._psi
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        # keep the contiguous copies alive across the ctypes calls (and for inspection)
        self.ops_c = dict(Dr=c(g.d_r), LIFT=c(g.lift), V=c(g.v), rx=c(g.r_x), Fscale=c(g.f_scale))
        self.ops_f = dict(Dr=c(gf.d_r), LIFT=c(gf.lift), V=c(gf.v), rx=c(gf.r_x), Fscale=c(gf.f_scale))
        self.P = c(g.prolongation_to(gf))
        o = self.ops_c
        self._check(self.lib.dgadj_set_operators(
            self._h, self.Np, self.K, _np_ptr(o["Dr"]), _np_ptr(o["LIFT"]), _np_ptr(o["V"]),
            _np_ptr(o["rx"]), _np_ptr(o["Fscale"])))
        self.has_adjoint = self.NpF <= 10        # N = 9 is forward-only (enriched space would be Np = 11)
        if self.has_adjoint:
            o = self.ops_f
            self._check(self.lib.dgadj_set_enriched(
                self._h, self.NpF, _np_ptr(o["Dr"]), _np_ptr(o["LIFT"]), _np_ptr(o["V"]),
                _np_ptr(o["rx"]), _np_ptr(o["Fscale"]), _np_ptr(self.P)))
            self.set_linear_functional(psi)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != _lib.OK:
            raise _lib.DgadjError(rc, self.lib.dgadj_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.dgadj_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _args(self, B, S, a, dt, t0, a_dev=None, dt_dev=None):
        return _lib.MarchArgs(B=B, S=S, reserved=0, t0=float(t0), a=float(a), dt=float(dt),
                              a_dev=a_dev, dt_dev=dt_dev)

    @staticmethod
    def _split_scalar(v):
        """(scalar, per-trajectory array or None)"""
        if np.isscalar(v):
            return float(v), None
        return 0.0, v

    def set_tuning(self, elems_per_thread=0, block_threads=0, grid_ctas=0):
        self._check(self.lib.dgadj_set_tuning(self._h, elems_per_thread, block_threads, grid_ctas))

    def plan(self, B, fused=True):
        ept, blk, tpc, grid = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        smem = C.c_int64()
        self._check(self.lib.dgadj_plan(self._h, B, int(fused), C.byref(ept), C.byref(blk), C.byref(tpc),
                                        C.byref(grid), C.byref(smem)))
        return dict(elems_per_thread=ept.value, block=blk.value, traj_per_cta=tpc.value, grid=grid.value,
                    smem_bytes=smem.value)

    def launch_count(self):
        return int(self.lib.dgadj_launch_count(self._h))

    def device_info(self):
        sm, mem, maj, mnr = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        self._check(self.lib.dgadj_device_info(self._h, C.byref(sm), C.byref(mem), C.byref(maj), C.byref(mnr)))
        return dict(sm_count=sm.value, total_mem=mem.value, cc=(maj.value, mnr.value))

    def measure_dfma_peak(self, seconds=0.5):
        tf, mhz = C.c_double(), C.c_double()
        self._check(self.lib.dgadj_measure_dfma_peak(self._h, float(seconds), C.byref(tf), C.byref(mhz)))
        return tf.value, mhz.value

    # ------------------------------------------------------------------ problem setup
    def set_linear_functional(self, psi=None):
        """J = int psi(x) u(x,T) dx  (psi = 1: 'J = int u', cf. getK in
        python/Main_finite_difference.py:153-155).  Nodal weights in both spaces."""
        self._psi = psi
        jw_c, jw_f = self.g.quad_weights(), self.gf.quad_weights()
        if psi is not None:
            jw_c, jw_f = jw_c * psi(self.g.x), jw_f * psi(self.gf.x)
        self.jw_c, self.jw_f = np.ascontiguousarray(jw_c), np.ascontiguousarray(jw_f)
        self._check(self.lib.dgadj_set_functional_weights(self._h, _np_ptr(self.jw_c), _np_ptr(self.jw_f)))

    def set_inflow_table(self, uin):
        uin = np.ascontiguousarray(uin, dtype=np.float64).ravel()
        self._check(self.lib.dgadj_set_inflow_table(self._h, uin.size, _np_ptr(uin)))

    def cfl_dt(self, FinalTime, CFL=0.75, speed=TWO_PI):
        """Time-step rule of utils/One_code.mlx: xmin = min|x(1,:)-x(2,:)|; dt = CFL/(2 pi) xmin;
        dt = dt/2; Nsteps = ceil(T/dt); dt = T/Nsteps.  Returns (dt, Nsteps)."""
        xmin = np.min(np.abs(self.g.x[0, :] - self.g.x[1, :]))
        dt = 0.5 * (CFL / speed * xmin)
        nsteps = int(math.ceil(FinalTime / dt))
        return FinalTime / nsteps, nsteps

    # ------------------------------------------------------------------ marches
    def _shape_u(self, u0, Np):
        if u0.ndim == 2:
            u0 = u0[None]
        if u0.ndim != 3 or u0.shape[1] != Np or u0.shape[2] != self.K:
            raise ValueError(f"expected (B, {Np}, {self.K}) or ({Np}, {self.K}), got {tuple(u0.shape)}")
        return u0

    def forward(self, u0, a, dt, S, t0=0.0, history=False):
        """Forward march (LSERK4 loop of utils/One_code.mlx + AdvecRHS1D, or explicit Euler).
        Returns uT, or (uT, hist[B, S+1, Np, K]) with history=True (the in-memory primal
        hand-off of matlab/adj_march.m:4)."""
        a_s, a_v = self._split_scalar(a)
        dt_s, dt_v = self._split_scalar(dt)
        if _is_host(u0):
            u0 = self._shape_u(np.ascontiguousarray(u0, dtype=np.float64), self.Np)
            B = u0.shape[0]
            a_v = None if a_v is None else np.ascontiguousarray(a_v, dtype=np.float64)
            dt_v = None if dt_v is None else np.ascontiguousarray(dt_v, dtype=np.float64)
            uT = np.empty_like(u0)
            hist = np.empty((B, S + 1, self.Np, self.K)) if history else None
            args = self._args(B, S, a_s, dt_s, t0)
            self._check(self.lib.dgadj_forward_host(self._h, C.byref(args), _np_ptr(a_v), _np_ptr(dt_v),
                                                    _np_ptr(u0), _np_ptr(uT), _np_ptr(hist)))
            return (uT, hist) if history else uT
        torch = _torch()
        u0 = self._shape_u(_as_device_tensor(u0), self.Np)
        B = u0.shape[0]
        a_v = None if a_v is None else _as_device_tensor(a_v)
        dt_v = None if dt_v is None else _as_device_tensor(dt_v)
        uT = torch.empty_like(u0)
        hist = torch.empty((B, S + 1, self.Np, self.K), dtype=torch.float64, device=u0.device) if history else None
        args = self._args(B, S, a_s, dt_s, t0, _ptr(a_v), _ptr(dt_v))
        self._check(self.lib.dgadj_forward(self._h, C.byref(args), _ptr(u0), _ptr(uT), _ptr(hist),
                                           C.c_void_p(0), self._stream()))
        return (uT, hist) if history else uT

    def fwd_adj(self, u0, a, dt, S, t0=0.0, want_uT=True, want_lam0=False, out=None, window=None, batch_chunk=0):
        """Fused forward + adjoint + indicator.  Returns dict(uT, J, eta[, lam0]):
        eta[B, K] is signed (consumers take abs, matlab/MAIN.m:51); lam0 = dJ/du0 in the
        enriched space (B, Np+1, K).  `out` may carry preallocated outputs (same keys).
        window (device tensors only; an int, or "auto" = only when the one-pass kernel's residual ring
        does not fit, with ceil(sqrt(S)) steps): march in windows with two-level
        checkpointing (`dgadj_fwd_adj_windowed`) -- for step counts whose residual ring does not fit
        the device; same results, about 1.3x the work."""
        a_s, a_v = self._split_scalar(a)
        dt_s, dt_v = self._split_scalar(dt)
        out = dict(out or {})
        if _is_host(u0):
            if window is not None:
                raise ValueError("window needs device tensors")
            u0 = self._shape_u(np.ascontiguousarray(u0, dtype=np.float64), self.Np)
            B = u0.shape[0]
            a_v = None if a_v is None else np.ascontiguousarray(a_v, dtype=np.float64)
            dt_v = None if dt_v is None else np.ascontiguousarray(dt_v, dtype=np.float64)
            uT = out.get("uT", np.empty_like(u0) if want_uT else None)
            J = out.get("J", np.empty(B))
            eta = out.get("eta", np.empty((B, self.K)))
            lam0 = out.get("lam0", np.empty((B, self.NpF, self.K)) if want_lam0 else None)
            args = self._args(B, S, a_s, dt_s, t0)
            self._check(self.lib.dgadj_fwd_adj_host(self._h, C.byref(args), _np_ptr(a_v), _np_ptr(dt_v),
                                                    _np_ptr(u0), _np_ptr(uT), _np_ptr(J), _np_ptr(lam0),
                                                    _np_ptr(eta)))
        else:
            torch = _torch()
            u0 = self._shape_u(_as_device_tensor(u0), self.Np)
            B = u0.shape[0]
            a_v = None if a_v is None else _as_device_tensor(a_v)
            dt_v = None if dt_v is None else _as_device_tensor(dt_v)
            kw = dict(dtype=torch.float64, device=u0.device)
            uT = out.get("uT", torch.empty_like(u0) if want_uT else None)
            J = out.get("J", torch.empty(B, **kw))
            eta = out.get("eta", torch.empty((B, self.K), **kw))
            lam0 = out.get("lam0", torch.empty((B, self.NpF, self.K), **kw) if want_lam0 else None)
            args = self._args(B, S, a_s, dt_s, t0, _ptr(a_v), _ptr(dt_v))
            def windowed(W):
                self._check(self.lib.dgadj_fwd_adj_windowed(self._h, C.byref(args), int(W), int(batch_chunk), _ptr(u0),
                                                            _ptr(uT), _ptr(J), _ptr(lam0), _ptr(eta), self._stream()))
            if window is not None and window != "auto":
                windowed(window)
            else:
                rc = self.lib.dgadj_fwd_adj(self._h, C.byref(args), _ptr(u0), _ptr(uT), _ptr(J), _ptr(lam0),
                                            _ptr(eta), self._stream())
                if rc == _lib.ERR_NOMEM and window == "auto":      # the residual ring does not fit: sqrt(S) windows
                    windowed(int(math.ceil(math.sqrt(S))))
                else:
                    self._check(rc)
        res = dict(J=J, eta=eta)
        if uT is not None:
            res["uT"] = uT
        if lam0 is not None:
            res["lam0"] = lam0
        return res

    def forward_checkpointed(self, u0, a, dt, S, t0=0.0):
        """Two-call path, step 1: forward march that also writes the checkpoints the adjoint
        consumes.  Device tensors only.  Returns (uT, ckpt)."""
        torch = _torch()
        a_s, a_v = self._split_scalar(a)
        dt_s, dt_v = self._split_scalar(dt)
        u0 = self._shape_u(_as_device_tensor(u0), self.Np)
        B = u0.shape[0]
        a_v = None if a_v is None else _as_device_tensor(a_v)
        dt_v = None if dt_v is None else _as_device_tensor(dt_v)
        nbytes = int(self.lib.dgadj_ckpt_bytes(self._h, B, S))
        if nbytes < 0:
            self._check(nbytes)
        ckpt = torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=u0.device)
        uT = torch.empty_like(u0)
        args = self._args(B, S, a_s, dt_s, t0, _ptr(a_v), _ptr(dt_v))
        self._check(self.lib.dgadj_forward(self._h, C.byref(args), _ptr(u0), _ptr(uT), C.c_void_p(0),
                                           _ptr(ckpt), self._stream()))
        return uT, ckpt

    def adjoint(self, uT, ckpt, a, dt, S, t0=0.0, want_lam0=True):
        """Two-call path, step 2: reverse-time adjoint march + indicator (conventions of
        matlab/adj_march.m:67-118 and errEst, python/Main_finite_difference.py:79-94)."""
        torch = _torch()
        a_s, a_v = self._split_scalar(a)
        dt_s, dt_v = self._split_scalar(dt)
        uT = self._shape_u(_as_device_tensor(uT), self.Np)
        B = uT.shape[0]
        a_v = None if a_v is None else _as_device_tensor(a_v)
        dt_v = None if dt_v is None else _as_device_tensor(dt_v)
        kw = dict(dtype=torch.float64, device=uT.device)
        J = torch.empty(B, **kw)
        eta = torch.empty((B, self.K), **kw)
        lam0 = torch.empty((B, self.NpF, self.K), **kw) if want_lam0 else None
        args = self._args(B, S, a_s, dt_s, t0, _ptr(a_v), _ptr(dt_v))
        self._check(self.lib.dgadj_adjoint(self._h, C.byref(args), _ptr(uT), _ptr(ckpt), _ptr(J), _ptr(lam0),
                                           _ptr(eta), self._stream()))
        res = dict(J=J, eta=eta)
        if lam0 is not None:
            res["lam0"] = lam0
        return res

    def rhs(self, u, timelocal, a, level=0):
        """rhsu = AdvecRHS1D(u, timelocal, a)  (utils/AdvecRHS1D.m:1-20); device tensors."""
        torch = _torch()
        Np = self.NpF if level else self.Np
        u = self._shape_u(_as_device_tensor(u), Np)
        a_s, a_v = self._split_scalar(a)
        a_v = None if a_v is None else _as_device_tensor(a_v)
        out = torch.empty_like(u)
        self._check(self.lib.dgadj_rhs(self._h, u.shape[0], level, _ptr(u), float(timelocal), a_s, _ptr(a_v),
                                       _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ refinement outputs
    def rank(self, eta, topk=1, want_order=True):
        """Refine flags / ranking of |eta| (matlab/MAIN.m:99,137; np.argmax,
        python/Main_finite_difference.py:337): order[B, K] int32 descending, ties by lowest
        index; flags[B, K] uint8 marks the topk elements.  Device tensors."""
        torch = _torch()
        eta = _as_device_tensor(eta)
        if eta.ndim == 1:
            eta = eta[None]
        B, K = eta.shape
        order = torch.empty((B, K), dtype=torch.int32, device=eta.device) if want_order else None
        flags = torch.empty((B, K), dtype=torch.uint8, device=eta.device)
        self._check(self.lib.dgadj_rank(self._h, B, K, _ptr(eta), int(topk), _ptr(order), _ptr(flags),
                                        self._stream()))
        return order, flags

    def ic_indicator(self, u0, u0f, lam0, eta):
        """eta[b,k] += lam0[b,:,k] . ((P u0)[:,k] - u0f[b,:,k]): the initial-data term of the indicator
        (`dgadj_ic_indicator`); u0 at the primal nodes [B,Np,K], u0f at the enriched nodes
        [B,NpF,K], lam0 from fwd_adj(want_lam0=True); eta updated in place.  Device tensors."""
        u0, u0f, lam0 = (_as_device_tensor(t).contiguous() for t in (u0, u0f, lam0))
        if not eta.is_contiguous():
            raise ValueError("eta must be contiguous (it is updated in place)")
        B = u0.shape[0]
        if u0f.shape != lam0.shape or lam0.shape[0] != B or eta.shape != (B, self.K):
            raise ValueError("shapes: u0 [B,Np,K], u0f / lam0 [B,NpF,K], eta [B,K]")
        self._check(self.lib.dgadj_ic_indicator(self._h, B, _ptr(u0), _ptr(u0f), _ptr(lam0), _ptr(eta), self._stream()))
        return eta

    def reduce_indicators(self, eta, J=None):
        """sums[K+4] = [sum_b |eta[b,k]| ..., sum|eta|, sum eta^2, max|eta|, sum J] in a fixed
        order (the per-rank partial of the batch-mean indicator,
        python/Main_variable_params.py:340).  Device tensors."""
        torch = _torch()
        eta = _as_device_tensor(eta)
        B, K = eta.shape
        J = None if J is None else _as_device_tensor(J)
        sums = torch.empty(K + 4, dtype=torch.float64, device=eta.device)
        self._check(self.lib.dgadj_reduce_indicators(self._h, B, K, _ptr(eta), _ptr(J), _ptr(sums), self._stream()))
        return sums

    def status(self, *fields):
        """status[B] int32 of a march's outputs (device tensors [B, ...]: uT, lam0, eta ...): bit 1
        (`_lib.STATUS_NON_FINITE`) = a NaN or Inf among the trajectory's values (`dgadj_march_status`)."""
        torch = _torch()
        out = None
        for f in fields:
            st = _lib.march_status(self.lib, self._h, torch, values=_as_device_tensor(f), stream=self._stream())
            out = st if out is None else (out | st)
        return out

    def set_element_orders(self, orders=None):
        """hp (Ns(k), matlab/MAIN.m:21,141, applied to the DG-in-space march; `dgadj_set_element_orders`): the
        polynomial order of every element, each in [1, N]; None returns to the uniform order N.  Fields keep the
        (B, N+1, K) layout on the order-N LGL nodes -- an element of lower order holds its polynomial's values there
        (it is L2-projected onto its own space on input).  The enriched space has one order more per element.
        Built for `forward` (without history), `fwd_adj` (also with `window=`) and `ic_indicator`; `adapt_advec(...,
        orders=)` runs the refinement loop with them."""
        if orders is None:
            self._check(self.lib.dgadj_set_element_orders(self._h, C.c_void_p(0)))
            self.orders = None
            return
        o = np.ascontiguousarray(orders, dtype=np.int32).ravel()
        if o.size != self.K:
            raise ValueError(f"one order per element: expected {self.K}, got {o.size}")
        nodes = np.ascontiguousarray(o + 1, dtype=np.int32)
        self._check(self.lib.dgadj_set_element_orders(self._h, _np_ptr(nodes)))
        self.orders = o.copy()

    def refine_shared(self, ind, v_x_dev, topk=1):
        """Split the `topk` elements with the largest |ind[k]| (lowest index on ties) of the mesh v_x_dev[:K+1]
        (a float64 CUDA tensor with room for K+topk+1 vertices) at their midpoints, on the device
        (`dgadj_refine_shared`; matlab/MAIN.m:137-141).  Returns the int32 CUDA tensor of the split elements."""
        torch = _torch()
        ind = _as_device_tensor(ind)
        refined = torch.empty(topk, dtype=torch.int32, device=ind.device)
        if v_x_dev.numel() < self.K + topk + 1:
            raise ValueError("v_x_dev has no room for the refined mesh")
        self._check(self.lib.dgadj_refine_shared(self._h, self.K, _ptr(ind), int(topk), _ptr(v_x_dev), _ptr(refined), self._stream()))
        return refined

    def reduce_indicator_blocks(self, eta, J=None, rows_per_block=None):
        """parts[nblk, K+4]: the sums of `reduce_indicators` over fixed blocks of `rows_per_block`
        trajectories (default sharding.REDUCE_BLOCK) -- the count-independent form that
        `sharding.allreduce_indicator_blocks` combines in global block order.  Device tensors."""
        from .sharding import REDUCE_BLOCK
        torch = _torch()
        eta = _as_device_tensor(eta)
        B, K = eta.shape
        R = int(rows_per_block or REDUCE_BLOCK)
        J = None if J is None else _as_device_tensor(J)
        nblk = (B + R - 1) // R
        parts = torch.empty((nblk, K + 4), dtype=torch.float64, device=eta.device)
        self._check(self.lib.dgadj_reduce_indicator_blocks(self._h, B, K, R, _ptr(eta), _ptr(J), _ptr(parts), self._stream()))
        return parts
