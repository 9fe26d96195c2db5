"""DG-in-time ODE path of the reference on the GPU: `dg_march`, `adj_march`
(matlab/dg_march.m, matlab/adj_march.m) and the refinement rule of matlab/MAIN.m:137-141,
batched over the initial value y0 on a shared mesh `times` (batch pattern of
python/Main_variable_params.py:330-339).

Host side (this file, once per mesh): the per-element constants -- fem_setup operators
(matlab/fem_setup.m:1-41 through `BaseGalerkin1D`), the polyfit/polyval interpolation of
dg_march.m:47-49 / adj_march.m:75-79 as matrices, including the reference's mirrored
quadrature interval (adj_march.m:72,78).  Device side (csrc/dgadj_tdg.cu):
everything that depends on the trajectory.  Orders must be uniform over the mesh.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .galerkin import BaseGalerkin1D


def _polyfit_matrix(x_from, deg, x_to):
    """polyval(polyfit(x_from, U, deg), x_to) as a matrix acting on U (quirk C-5)."""
    n = len(x_from)
    M = np.zeros((len(x_to), n))
    eye = np.eye(n)
    for j in range(n):
        M[:, j] = np.polyval(np.polyfit(x_from, eye[j], deg), x_to)
    return M


class TimeDG:
    def __init__(self, linear=False, device=0, tol=1e-7, maxit=500):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.linear, self.device, self.tol, self.maxit = bool(linear), device, float(tol), int(maxit)
        self._cache = {}                        # per-element constants by (kind, order, element end points)
        cfg = _lib.Config(device=device, N=1, K=1, bc=1, inflow=0, functional=0, scheme=0, reserved=0, alpha=0.0)
        self._h = C.c_void_p(0)
        rc = self.lib.dgadj_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.OK:
            self._h = C.c_void_p(0)
            raise _lib.DgadjError(rc, "dgadj_create failed (an sm_100 device is required; there is no CPU path)")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.dgadj_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _lib.OK:
            raise _lib.DgadjError(rc, self.lib.dgadj_last_error(self._h).decode())

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _uniform(Ns):
        Ns = np.asarray(Ns).astype(int).ravel()
        if not np.all(Ns == Ns[0]):
            raise _lib.DgadjError(_lib.ERR_UNSUPPORTED, "mixed orders over the mesh are not supported")
        return int(Ns[0])

    # ------------------------------------------------------------------ constants
    def march_constants(self, N, times):
        """Per-element block A | Iq | Phi | w | hk of dg_march.m (layout: csrc/dgadj_tdg.cu)."""
        Ks = len(times) - 1
        n_gq = 1 if self.linear else 30 * N                       # dg_march.m:13 / :29
        blocks, nodes = [], []
        nq = 0
        for k in range(Ks):
            key = ("m", N, float(times[k]), float(times[k + 1]))
            if key in self._cache:
                blk, x, nq = self._cache[key]
                blocks.append(blk); nodes.append(x)
                continue
            g = BaseGalerkin1D(n=N, k=1, domain=(times[k], times[k + 1]), n_gq=n_gq)
            x = g.x[:, 0]
            hk = x[-1] - x[0]                                     # :14 / :30
            Minv = g.mass
            S = Minv @ g.d_r                                      # :16 / :53
            Np = g.n_p
            Bm = np.zeros((Np, Np))
            if self.linear:
                Bm[-1, -1] = 1.0                                  # :17
                A = -S.T + Bm - hk / 2 * Minv                     # :15,:18
                parts = [A.ravel(), [hk]]
                nq = 0
            else:
                Bm[-1, -1] = -1.0                                 # :54
                A = S.T + Bm                                      # :57
                x_interp = x[0] + (1 + g.r) * hk / 2              # :48
                Iq = _polyfit_matrix(x, N, x_interp)              # :47-49
                nq = g.n_r
                parts = [A.ravel(), Iq.ravel(), g.phi.ravel(), g.w, [hk]]
            blocks.append(np.concatenate([np.asarray(p, dtype=np.float64) for p in parts]))
            nodes.append(x)
            self._cache[key] = (blocks[-1], x, nq)
        return np.ascontiguousarray(np.concatenate(blocks)), nodes, nq

    def adjoint_constants(self, Na, t1):
        """Per-element block A0 | f1 | A2 | Ix | Iq | Phi | w | hk of adj_march.m."""
        blocks, nodes = [], []
        nq = 0
        for tk in t1:
            tk = np.asarray(tk, dtype=np.float64)
            key = ("a", Na, tk.tobytes())
            if key in self._cache:
                blk, x, nq = self._cache[key]
                blocks.append(blk); nodes.append(x)
                continue
            g = BaseGalerkin1D(n=Na, k=1, domain=(tk[0], tk[-1]), n_gq=1 if self.linear else 2 * Na)  # :17 / :71
            x = g.x[:, 0]
            hk = x[0] - x[-1]                                     # :18 / :72  negative (quirk C-3)
            Minv = g.mass
            S = Minv @ g.d_r
            Np = g.n_p
            deg = len(tk) - 1                                     # :36 / :75
            Ix = _polyfit_matrix(tk, deg, x)
            M = hk / 2 * Minv
            f1 = M @ np.ones(Np)                                  # :28 / :96
            if self.linear:
                m = np.zeros((Np, Np)); m[0, 0] = -1.0            # :21
                A0 = -S.T + m - M                                 # :22
                m2 = np.zeros((Np, Np)); m2[-1, -1] = 1.0         # :40
                A2 = -S.T + m2 + M                                # :41
                parts = [A0.ravel(), f1, A2.ravel(), Ix.ravel(), [hk]]
                nq = 0
            else:
                Bm = np.zeros((Np, Np)); Bm[0, 0] = -1.0          # :85
                A0 = -S.T + Bm                                    # :86 without M_v (state dependent)
                B2 = np.zeros((Np, Np)); B2[-1, -1] = -1.0        # :107
                A2 = -S.T - B2                                    # :115
                r_interp = tk[0] + (1 + g.r) * hk / 2             # :78 (mirrored interval)
                Iq = _polyfit_matrix(tk, deg, r_interp)
                nq = g.n_r
                parts = [A0.ravel(), f1, A2.ravel(), Ix.ravel(), Iq.ravel(), g.phi.ravel(), g.w, [hk]]
            blocks.append(np.concatenate([np.asarray(p, dtype=np.float64) for p in parts]))
            nodes.append(x)
            self._cache[key] = (blocks[-1], x, nq)
        return np.ascontiguousarray(np.concatenate(blocks)), nodes, nq

    # ------------------------------------------------------------------ reference-named entry points
    def dg_march(self, Ns, Ks, times, y0, x_true=None, u_true=None):
        """[t, y] = dg_march(Ns, Ks, times, y0, x_true, u_true)  (matlab/dg_march.m:1).
        y0: float64 CUDA tensor [B].  Returns (t, y, its): t = list of Ks node arrays,
        y[B, Ks, Np], its[B, Ks] Newton iteration counts (the reference prints them, :70)."""
        torch = self.torch
        N = self._uniform(Ns)
        y0 = y0.contiguous().view(-1)
        B = y0.numel()
        consts, nodes, nq = self.march_constants(N, np.asarray(times, dtype=np.float64))
        y = torch.empty((B, Ks, N + 1), dtype=torch.float64, device=y0.device)
        its = torch.empty((B, Ks), dtype=torch.int32, device=y0.device)
        self._check(self.lib.dgadj_tdg_march(self._h, B, Ks, N + 1, nq, int(self.linear), self.tol, self.maxit,
                                             C.c_void_p(consts.ctypes.data), C.c_void_p(y0.data_ptr()),
                                             C.c_void_p(y.data_ptr()), C.c_void_p(its.data_ptr()), self._stream()))
        return nodes, y, its

    def adj_march(self, Ns, Ks, times, y1, t1, y0=1.0):
        """[t, v, err] = adj_march(Ns, Ks, times)  (matlab/adj_march.m:1); the primal the
        reference reads from globals (`y1`, `t1`, :4) is passed explicitly.  Ns = adjoint orders
        (matlab/MAIN.m:34 passes Ns+1).  Returns (t, v[B, Ks, Na+1], err[B, Ks]) -- err signed."""
        torch = self.torch
        Na = self._uniform(Ns)
        Npp = y1.shape[2]
        if Na + 1 != Npp + 1:
            raise _lib.DgadjError(_lib.ERR_UNSUPPORTED, "adjoint order must be primal order + 1 (matlab/MAIN.m:34)")
        B = y1.shape[0]
        consts, nodes, nq = self.adjoint_constants(Na, t1)
        v = torch.empty((B, Ks, Na + 1), dtype=torch.float64, device=y1.device)
        err = torch.empty((B, Ks), dtype=torch.float64, device=y1.device)
        self._check(self.lib.dgadj_tdg_adjoint(self._h, B, Ks, Npp, nq, int(self.linear), float(y0),
                                               C.c_void_p(consts.ctypes.data), C.c_void_p(y1.contiguous().data_ptr()),
                                               C.c_void_p(v.data_ptr()), C.c_void_p(err.data_ptr()), self._stream()))
        return nodes, v, err


def refine(times, Ns, err_mean, n):
    """matlab/MAIN.m:137-141: refine the element with the largest |err| (lowest index on ties,
    quirk C-10) by midpoint insertion; a new order entry n is appended."""
    err_mean = np.asarray(err_mean)
    ref_i = int(np.argmax(np.abs(err_mean)))
    times = np.asarray(times, dtype=np.float64)
    times = np.insert(times, ref_i + 1, 0.5 * (times[ref_i] + times[ref_i + 1]))
    return times, np.append(np.asarray(Ns), n), ref_i
