"""Inviscid Burgers DG march with the reference's slope limiter fused into every RK stage
(BASELINE config 3).  Limiter: utils/SlopeLimitN.m / SlopeLimitLin.m / minmod.m; the Burgers
right-hand side is build-specified (the reference has none).  The march
writes the forward checkpoints a discrete adjoint needs (per-step states, per-stage limiter
flags and wave speeds)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .galerkin import BaseGalerkin1D


class BurgersDG1D:
    def __init__(self, N, K=None, domain=(-1.0, 1.0), v_x=None, bc="periodic", device=0):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.g = g = BaseGalerkin1D(n=N, k=K, domain=domain, v_x=v_x)
        self.N, self.Np, self.K, self.device, self.bc = N, N + 1, g.k, device, bc
        if bc not in ("periodic", "free"):
            raise ValueError("bc must be 'periodic' or 'free' (zero-jump ends)")
        cfg = _lib.Config(device=device, N=N, K=self.K, bc=1 if bc == "periodic" else 0, inflow=0, functional=0,
                          scheme=0, reserved=0, alpha=0.0)
        self._h = C.c_void_p(0)
        rc = self.lib.dgadj_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.OK:
            self._h = C.c_void_p(0)
            raise _lib.DgadjError(rc, "dgadj_create failed (an sm_100 device is required; there is no CPU path)")
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.ops = dict(Dr=c(g.d_r), LIFT=c(g.lift), Mref=c(g.mass), rx=c(g.r_x), Fscale=c(g.f_scale),
                        V=c(g.v), invV=c(g.inv_v), x=c(g.x))
        o = self.ops
        p = lambda a: C.c_void_p(a.ctypes.data)
        self._check(self.lib.dgadj_set_operators(self._h, self.Np, self.K, p(o["Dr"]), p(o["LIFT"]), p(o["V"]),
                                                 p(o["rx"]), p(o["Fscale"])))

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

    def stable_dt(self, umax, cfl=0.25):
        """dt = cfl * min node spacing / max wave speed."""
        xmin = np.min(np.abs(self.g.x[0, :] - self.g.x[1, :]))
        return cfl * xmin / umax

    LIMIT_CODES = {False: 0, None: 0, 0: 0, True: 1, 1: 1, "N": 1, "1": 2, "Pi1": 2, 2: 2}

    def forward(self, u0, dt, S, limit=True, history=False, checkpoints=False, tvb_M=0.0):
        """u0: float64 CUDA tensor [B, Np, K]; dt scalar or CUDA tensor [B].
        limit: True / "N" = SlopeLimitN after every stage (utils/SlopeLimitN.m), "1" = SlopeLimit1 (every
        cell, utils/SlopeLimit1.m), False = none; tvb_M = the M of the TVB minmod (utils/minmodB.m), 0 = minmod.
        Returns dict(uT[, hist[B,S+1,Np,K]][, lim[B,S,K] int16, lim0[B,K] uint8, amax[B,S,5] int32,
        maxvel[B,S,5]]).  `checkpoints=True` writes everything `adjoint` consumes (incl. hist)."""
        torch = self.torch
        if not (isinstance(u0, torch.Tensor) and u0.is_cuda and u0.dtype == torch.float64):
            raise TypeError("u0 must be a float64 CUDA tensor")
        if u0.ndim == 2:
            u0 = u0[None]
        if u0.shape[1:] != (self.Np, self.K):
            raise ValueError(f"expected (B, {self.Np}, {self.K}), got {tuple(u0.shape)}")
        u0 = u0.contiguous()
        B = u0.shape[0]
        dt_s, dt_v = (float(dt), None) if np.isscalar(dt) else (0.0, dt.contiguous())
        kw = dict(dtype=torch.float64, device=u0.device)
        out = dict(uT=torch.empty_like(u0))
        if history or checkpoints:
            out["hist"] = torch.empty((B, S + 1, self.Np, self.K), **kw)
        if checkpoints:
            out["lim"] = torch.zeros((B, S, self.K), dtype=torch.int16, device=u0.device)
            out["lim0"] = torch.zeros((B, self.K), dtype=torch.uint8, device=u0.device)
            out["amax"] = torch.zeros((B, S, 5), dtype=torch.int32, device=u0.device)
            out["maxvel"] = torch.empty((B, S, 5), **kw)
        ptr = lambda k: C.c_void_p(out[k].data_ptr()) if k in out else C.c_void_p(0)
        o = self.ops
        p = lambda a: C.c_void_p(a.ctypes.data)
        self._check(self.lib.dgadj_burgers_forward(
            self._h, B, S, dt_s, C.c_void_p(dt_v.data_ptr()) if dt_v is not None else C.c_void_p(0), self.LIMIT_CODES[limit],
            float(tvb_M), p(o["invV"]), p(o["V"]), p(o["x"]), C.c_void_p(u0.data_ptr()), ptr("uT"), ptr("hist"), ptr("lim"),
            ptr("lim0"), ptr("amax"), ptr("maxvel"), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        out["dt"], out["S"] = dt, S
        return out

    def adjoint(self, fwd, psi=None):
        """Discrete adjoint of a `forward(..., checkpoints=True)` run for J = int psi(x) u(x,T) dx
        (psi = 1 by default): returns dict(lam0[B,Np,K] = dJ/du0, J[B]).  The limiter and the
        max|u| of the Lax-Friedrichs flux are transposed on the branches the forward run took."""
        torch = self.torch
        for k in ("hist", "lim", "lim0", "amax", "maxvel"):
            if k not in fwd:
                raise ValueError("adjoint needs a forward run with checkpoints=True")
        jw = self.g.quad_weights()
        if psi is not None:
            jw = jw * psi(self.g.x)
        jw = np.ascontiguousarray(jw, dtype=np.float64)
        B, S = fwd["hist"].shape[0], fwd["S"]
        dt = fwd["dt"]
        dt_s, dt_v = (float(dt), None) if np.isscalar(dt) else (0.0, dt.contiguous())
        kw = dict(dtype=torch.float64, device=fwd["hist"].device)
        lam0 = torch.empty((B, self.Np, self.K), **kw)
        J = torch.empty(B, **kw)
        o = self.ops
        p = lambda a: C.c_void_p(a.ctypes.data)
        d = lambda t: C.c_void_p(t.data_ptr())
        self._check(self.lib.dgadj_burgers_adjoint(
            self._h, B, S, dt_s, d(dt_v) if dt_v is not None else C.c_void_p(0), p(o["invV"]), p(o["V"]), p(o["x"]),
            p(jw), d(fwd["hist"]), d(fwd["lim"]), d(fwd["lim0"]), d(fwd["amax"]), d(fwd["maxvel"]), d(lam0), d(J),
            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return dict(lam0=lam0, J=J)

    def _enriched(self):
        """Operators of the enriched space (order N+1) for the indicator, set on first use."""
        if getattr(self, "gf", None) is None:
            if self.N + 2 > 10:
                raise ValueError("the indicator needs N <= 8 (the enriched space is order N+1)")
            g = self.g
            self.gf = gf = BaseGalerkin1D(n=self.N + 1, k=self.K, domain=g.domain, v_x=g.v_x)
            c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
            self.ops_f = dict(Dr=c(gf.d_r), LIFT=c(gf.lift), rx=c(gf.r_x), Fscale=c(gf.f_scale), V=c(gf.v),
                              invV=c(gf.inv_v), x=c(gf.x))
            self.P = c(g.prolongation_to(gf))
            o = self.ops_f
            p = lambda a: C.c_void_p(a.ctypes.data)
            self._check(self.lib.dgadj_set_enriched(self._h, self.Np + 1, p(o["Dr"]), p(o["LIFT"]), p(o["V"]), p(o["rx"]),
                                                    p(o["Fscale"]), p(self.P)))
        return self.gf

    def plan(self, B, indicator=True):
        """Launch shape of the fused kernel: dict(elems_per_thread, block, grid, smem_bytes, ring_bytes_per_step)."""
        if indicator:
            self._enriched()
        e, b, g = C.c_int32(), C.c_int32(), C.c_int32()
        sm, rb = C.c_int64(), C.c_int64()
        self._check(self.lib.dgadj_burgers_plan(self._h, int(B), int(bool(indicator)), C.byref(e), C.byref(b), C.byref(g),
                                                C.byref(sm), C.byref(rb)))
        return dict(elems_per_thread=e.value, block=b.value, grid=g.value, smem_bytes=sm.value, ring_bytes_per_step=rb.value)

    def set_tuning(self, elems_per_thread=0, grid_ctas=0):
        self._check(self.lib.dgadj_set_tuning(self._h, elems_per_thread, 0, grid_ctas))

    def fwd_adj(self, u0, dt, S, limit=True, indicator=True, psi=None, tvb_M=0.0, want_uT=True, want_lam0=True):
        """BASELINE config 3 in one call (`dgadj_burgers_fwd_adj`): the limited march, its discrete adjoint on
        the frozen limiter / minmod / argmax branches and, with indicator=True, the per-element error indicator
        (adjoint one order higher, matlab/MAIN.m:34; err(k) = v_k' * residual, matlab/adj_march.m:103-117).
        Forward states go to a per-CTA ring on the device (no [B, S+1, Np, K] history), so marches past shock
        formation fit at the full batch.  J = int psi(x) u(x,T) dx (psi = 1 by default).
        Returns dict(uT, J[B], lam0, nlim[B, 2] int32 (limiter activations of the march / of the steps the adjoint
        phase takes again), status[B] int32[, eta[B, K]]):
          indicator=False: lam0[B, Np, K] = dJ/du0 of the march (what `adjoint` gives);
          indicator=True:  lam0[B, Np+1, K] = the enriched adjoint at t = 0, eta signed (consumers take abs)."""
        torch = self.torch
        if not (isinstance(u0, torch.Tensor) and u0.is_cuda and u0.dtype == torch.float64):
            raise TypeError("u0 must be a float64 CUDA tensor")
        if u0.ndim == 2:
            u0 = u0[None]
        if u0.shape[1:] != (self.Np, self.K):
            raise ValueError(f"expected (B, {self.Np}, {self.K}), got {tuple(u0.shape)}")
        u0 = u0.contiguous()
        B = u0.shape[0]
        dt_s, dt_v = (float(dt), None) if np.isscalar(dt) else (0.0, dt.contiguous())
        jw = self.g.quad_weights()
        if psi is not None:
            jw = jw * psi(self.g.x)
        jw = np.ascontiguousarray(jw, dtype=np.float64)
        o = self.ops
        p = lambda a: C.c_void_p(a.ctypes.data)
        a = _lib.BurgersArgs(B=B, S=int(S), limit=self.LIMIT_CODES[limit], indicator=int(bool(indicator)), reserved=0,
                             dt=dt_s, dt_dev=C.c_void_p(dt_v.data_ptr()) if dt_v is not None else None, tvb_M=float(tvb_M),
                             invV_host=p(o["invV"]), V_host=p(o["V"]), x_host=p(o["x"]), jw_host=p(jw))
        keep = [jw]
        NpX = self.Np
        if indicator:
            gf = self._enriched()
            jwf = gf.quad_weights()
            if psi is not None:
                jwf = jwf * psi(gf.x)
            jwf = np.ascontiguousarray(jwf, dtype=np.float64)
            keep.append(jwf)
            of = self.ops_f
            a.invVF_host, a.VF_host, a.xF_host, a.jwF_host = p(of["invV"]), p(of["V"]), p(of["x"]), p(jwf)
            NpX = self.Np + 1
        kw = dict(dtype=torch.float64, device=u0.device)
        out = dict(J=torch.empty(B, **kw), nlim=torch.empty((B, 2), dtype=torch.int32, device=u0.device),
                   status=torch.empty(B, dtype=torch.int32, device=u0.device))
        if want_uT:
            out["uT"] = torch.empty_like(u0)
        if want_lam0:
            out["lam0"] = torch.empty((B, NpX, self.K), **kw)
        if indicator:
            out["eta"] = torch.empty((B, self.K), **kw)
        d = lambda k: C.c_void_p(out[k].data_ptr()) if k in out else C.c_void_p(0)
        self._check(self.lib.dgadj_burgers_fwd_adj(
            self._h, C.byref(a), C.c_void_p(u0.data_ptr()), d("uT"), d("J"), d("lam0"), d("eta"), d("nlim"), d("status"),
            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out

    def slope_limit(self, u, kind="N", tvb_M=0.0):
        """ulimit = SlopeLimitN(u) (utils/SlopeLimitN.m:1) or, kind="1", SlopeLimit1(u)
        (utils/SlopeLimit1.m:1): one limiter pass, no time steps; tvb_M > 0 uses minmodB."""
        return self.forward(u, 0.0, 0, limit=kind, tvb_M=tvb_M)["uT"]


def decode_limiter_record(lim):
    """Packed per-step limiter record [B, S, K] (int16) -> (flags, branches), each [B, S, 5, K]."""
    l = lim.int() & 0xFFFF
    st = l.new_tensor(range(5)).view(1, 1, 5, 1)
    flags = ((l[:, :, None, :] >> st) & 1).bool()
    branches = (l[:, :, None, :] >> (5 + 2 * st)) & 3
    return flags, branches
