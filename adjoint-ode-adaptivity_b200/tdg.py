"""DG-in-time ODE path of the reference on the GPU: `dg_march`, `adj_march`
(matlab/dg_march.m, matlab/adj_march.m) and the refinement rule of matlab/MAIN.m:137-141,
batched over the initial value y0 on a shared mesh `times` (batch pattern of
python/Main_variable_params.py:330-339).

Host side (this file, once per mesh): the per-element constants -- fem_setup operators
(matlab/fem_setup.m:1-41 through `BaseGalerkin1D`), the polyfit/polyval interpolation of
dg_march.m:47-49 / adj_march.m:75-79 as matrices, including the reference's mirrored
quadrature interval (adj_march.m:72,78).  Device side (csrc/dgadj_tdg.cu):
everything that depends on the trajectory.  Orders may differ between elements (Ns(k),
matlab/MAIN.m:21,141): the per-element blocks are padded to the mesh maxima.
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


# utils/Globals1D.m:37-42: left Radau points as the reference tabulates them (six-digit decimals
# for m = 4, 5 -- kept as written)
RADAU = {
    1: np.array([-1.0]),
    2: np.array([-1.0, 1.0 / 3.0]),
    3: np.array([-1.0, (1.0 - np.sqrt(6.0)) / 5.0, (1.0 + np.sqrt(6.0)) / 5.0]),
    4: np.array([-1.0, -0.575319, 0.181066, 0.822824]),
    5: np.array([-1.0, -0.72048, -0.167181, 0.446314, 0.885792]),
}


class TimeDG:
    def __init__(self, linear=False, device=0, tol=1e-7, maxit=500, quirks=True):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.linear, self.device, self.tol, self.maxit = bool(linear), device, float(tol), int(maxit)
        # quirks=True: the reference bug for bug.  quirks=False switches off SURVEY quirk C-3 only: the
        # adjoint linearises about the primal at quadrature points INSIDE the element instead of the
        # mirrored interval of adj_march.m:72,78 -- with it the element indicators sum to the error of J
        # (tools/cfg5_report.py); everything else is unchanged.
        self.quirks = bool(quirks)
        self._cache = {}                        # per-element constants by (kind, order, element end points)
        self._mesh_cache = {}                   # padded whole-mesh blocks by (kind, orders, mesh)
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
    def _orders(Ns, Ks):
        Ns = np.asarray(Ns).astype(int).ravel()
        if Ns.size == 1:
            Ns = np.repeat(Ns, Ks)
        if Ns.size != Ks or Ns.min() < 1:
            raise _lib.DgadjError(_lib.ERR_INVALID, "Ns must hold one order >= 1 per element")
        return Ns

    @staticmethod
    def _pad(M, rows, cols, identity=False):
        """Pad a per-element matrix to the mesh-wide (rows, cols): zeros, or identity rows for a
        system matrix, so that the padded unknowns are exactly 0 (layout note in csrc/dgadj_tdg.cu)."""
        M = np.atleast_2d(np.asarray(M, dtype=np.float64))
        out = np.zeros((rows, cols))
        out[:M.shape[0], :M.shape[1]] = M
        if identity:
            for i in range(M.shape[0], rows):
                out[i, i] = 1.0
        return out.ravel()

    # ------------------------------------------------------------------ constants
    def _march_element(self, N, t0, t1):
        key = ("m", N, float(t0), float(t1))
        if key not in self._cache:
            n_gq = 1 if self.linear else 30 * N                   # dg_march.m:13 / :29
            g = BaseGalerkin1D(n=N, k=1, domain=(t0, t1), n_gq=n_gq)
            x = g.x[:, 0]
            hk = x[-1] - x[0]                                     # :14 / :30
            Minv = g.mass
            S = Minv @ g.d_r                                      # :16 / :53
            Np = g.n_p
            Bm = np.zeros((Np, Np))
            el = dict(x=x, hk=hk, Np=Np, nq=0)
            if self.linear:
                Bm[-1, -1] = 1.0                                  # :17
                el["A"] = -S.T + Bm - hk / 2 * Minv               # :15,:18
            else:
                Bm[-1, -1] = -1.0                                 # :54
                el["A"] = S.T + Bm                                # :57
                x_interp = x[0] + (1 + g.r) * hk / 2              # :48
                el["Iq"] = _polyfit_matrix(x, N, x_interp)        # :47-49
                el["Phi"], el["w"], el["nq"] = g.phi, g.w, g.n_r
            self._cache[key] = el
        return self._cache[key]

    def march_constants(self, Ns, times):
        """Per-element block A | Iq | Phi | w | hk | np_k of dg_march.m, padded to the largest
        order / quadrature of the mesh (layout: csrc/dgadj_tdg.cu).  Returns (blocks, node
        arrays, Np_max, nq_max)."""
        Ks = len(times) - 1
        Ns = self._orders(Ns, Ks)
        mkey = ("M", Ns.tobytes(), np.asarray(times, dtype=np.float64).tobytes())
        if mkey in self._mesh_cache:
            return self._mesh_cache[mkey]
        els = [self._march_element(int(Ns[k]), times[k], times[k + 1]) for k in range(Ks)]
        NP = max(e["Np"] for e in els)
        nq = max(e["nq"] for e in els)
        blocks = []
        for e in els:
            parts = [self._pad(e["A"], NP, NP, identity=True)]
            if not self.linear:
                parts += [self._pad(e["Iq"], nq, NP), self._pad(e["Phi"], nq, NP), self._pad(e["w"], 1, nq)]
            parts += [[e["hk"], float(e["Np"])]]
            blocks.append(np.concatenate([np.asarray(p, dtype=np.float64) for p in parts]))
        out = (np.ascontiguousarray(np.concatenate(blocks)), [e["x"] for e in els], NP, nq)
        self._remember(mkey, out)
        return out

    def _remember(self, key, value):
        if len(self._mesh_cache) >= 8:          # whole-mesh blocks of the last few meshes only
            self._mesh_cache.pop(next(iter(self._mesh_cache)))
        self._mesh_cache[key] = value

    def _adjoint_element(self, Na, tk):
        key = ("a", Na, self.quirks, tk.tobytes())
        if key not in self._cache:
            g = BaseGalerkin1D(n=Na, k=1, domain=(tk[0], tk[-1]), n_gq=1 if self.linear else 2 * Na)  # :17 / :71
            x = g.x[:, 0]
            hk = x[0] - x[-1]                                     # :18 / :72  negative (quirk C-3)
            Minv = g.mass
            S = Minv @ g.d_r
            Np = g.n_p
            deg = len(tk) - 1                                     # :36 / :75
            M = hk / 2 * Minv
            el = dict(x=x, hk=hk, Na=Np, Npp=len(tk), nq=0, Ix=_polyfit_matrix(tk, deg, x),
                      f1=M @ np.ones(Np))                         # :28 / :96
            if self.linear:
                m = np.zeros((Np, Np)); m[0, 0] = -1.0            # :21
                el["A0"] = -S.T + m - M                           # :22
                m2 = np.zeros((Np, Np)); m2[-1, -1] = 1.0         # :40
                el["A2"] = -S.T + m2 + M                          # :41
            else:
                Bm = np.zeros((Np, Np)); Bm[0, 0] = -1.0          # :85
                el["A0"] = -S.T + Bm                              # :86 without M_v (state dependent)
                B2 = np.zeros((Np, Np)); B2[-1, -1] = -1.0        # :107
                el["A2"] = -S.T - B2                              # :115
                r_interp = tk[0] + (1 + g.r) * (hk if self.quirks else -hk) / 2    # :78 (mirrored interval: C-3)
                el["Iq"] = _polyfit_matrix(tk, deg, r_interp)
                el["Phi"], el["w"], el["nq"] = g.phi, g.w, g.n_r
            self._cache[key] = el
        return self._cache[key]

    def adjoint_constants(self, Nas, t1):
        """Per-element block A0 | f1 | A2 | Ix | Iq | Phi | w | hk | na_k | last_{k-1} of
        adj_march.m, padded to the mesh maxima.  Returns (blocks, node arrays, Npp_max, nq_max)."""
        Ks = len(t1)
        Nas = self._orders(Nas, Ks)
        mkey = ("A", self.quirks, Nas.tobytes(), b"".join(np.asarray(t, dtype=np.float64).tobytes() for t in t1))
        if mkey in self._mesh_cache:
            return self._mesh_cache[mkey]
        els = [self._adjoint_element(int(Nas[k]), np.asarray(t1[k], dtype=np.float64)) for k in range(Ks)]
        for e in els:
            if e["Na"] != e["Npp"] + 1:
                raise _lib.DgadjError(_lib.ERR_UNSUPPORTED,
                                      "adjoint order must be primal order + 1 on every element (matlab/MAIN.m:34)")
        NPP = max(e["Npp"] for e in els)
        NA = NPP + 1
        nq = max(e["nq"] for e in els)
        blocks = []
        for k, e in enumerate(els):
            parts = [self._pad(e["A0"], NA, NA, identity=True), self._pad(e["f1"], 1, NA), self._pad(e["A2"], NA, NA),
                     self._pad(e["Ix"], NA, NPP)]
            if not self.linear:
                parts += [self._pad(e["Iq"], nq, NPP), self._pad(e["Phi"], nq, NA), self._pad(e["w"], 1, nq)]
            parts += [[e["hk"], float(e["Na"]), float(els[k - 1]["Npp"] - 1 if k > 0 else 0)]]
            blocks.append(np.concatenate([np.asarray(p, dtype=np.float64) for p in parts]))
        out = (np.ascontiguousarray(np.concatenate(blocks)), [e["x"] for e in els], NPP, nq)
        self._remember(mkey, out)
        return out

    def _y0_args(self, y0, B):
        """(scalar, device pointer) for the initial value the first element's residual is measured
        against: the reference hard-codes y0 = 1 (adj_march.m:9); a float64 CUDA tensor [B] gives every
        trajectory of a batch its own."""
        if isinstance(y0, self.torch.Tensor):
            y0 = y0.contiguous().view(-1)
            if y0.numel() != B or y0.dtype != self.torch.float64 or not y0.is_cuda:
                raise _lib.DgadjError(_lib.ERR_INVALID, "y0 must be a float64 CUDA tensor [B]")
            self._y0_keep = y0
            return 0.0, C.c_void_p(y0.data_ptr())
        return float(y0), C.c_void_p(0)

    # ------------------------------------------------------------------ reference-named entry points
    def dg_march(self, Ns, Ks, times, y0, x_true=None, u_true=None):
        """[t, y] = dg_march(Ns, Ks, times, y0, x_true, u_true)  (matlab/dg_march.m:1).
        y0: float64 CUDA tensor [B].  Ns: one order per element (or a scalar); orders may differ
        between elements (Ns(k), matlab/MAIN.m:21,141).  Returns (t, y, its): t = list of Ks node
        arrays, y[B, Ks, Np_max] (element k holds len(t[k]) values, zero beyond), its[B, Ks]
        Newton iteration counts (the reference prints them, :70)."""
        torch = self.torch
        y0 = y0.contiguous().view(-1)
        B = y0.numel()
        consts, nodes, NP, nq = self.march_constants(Ns, np.asarray(times, dtype=np.float64))
        y = torch.empty((B, Ks, NP), dtype=torch.float64, device=y0.device)
        its = torch.empty((B, Ks), dtype=torch.int32, device=y0.device)
        self._check(self.lib.dgadj_tdg_march(self._h, B, Ks, NP, nq, int(self.linear), self.tol, self.maxit,
                                             C.c_void_p(consts.ctypes.data), C.c_void_p(y0.data_ptr()),
                                             C.c_void_p(y.data_ptr()), C.c_void_p(its.data_ptr()), self._stream()))
        # per-trajectory status word (the reference prints "not converged", dg_march.m:69-73): kept for `status()`
        self._last = (y, its)
        return nodes, y, its

    def status(self, y=None, its=None):
        """status[B] int32 of a march (default: the last `dg_march`): bit 0 = a Newton solve stopped at maxit + 1
        iterations without meeting the tolerance (dg_march.m:69-73), bit 1 = a non-finite value in y."""
        if y is None and its is None:
            y, its = self._last
        return _lib.march_status(self.lib, self._h, self.torch, values=y, its=its, maxit=self.maxit, stream=self._stream())

    def adj_march(self, Ns, Ks, times, y1, t1, y0=1.0):
        """[t, v, err] = adj_march(Ns, Ks, times)  (matlab/adj_march.m:1); the primal the
        reference reads from globals (`y1`, `t1`, :4) is passed explicitly as dg_march returned it.
        Ns = adjoint orders, primal order + 1 on every element (matlab/MAIN.m:34 passes Ns+1).
        y0 = the initial value(s): the reference's hard-coded 1 (adj_march.m:9) or a CUDA tensor [B].
        Returns (t, v[B, Ks, Na_max], err[B, Ks]) -- err signed, v zero beyond len(t[k])."""
        torch = self.torch
        B = y1.shape[0]
        consts, nodes, NPP, nq = self.adjoint_constants(Ns, t1)
        if y1.shape[2] != NPP:
            raise _lib.DgadjError(_lib.ERR_INVALID, "y1 must be the [B, Ks, Np_max] array dg_march returned")
        v = torch.empty((B, Ks, NPP + 1), dtype=torch.float64, device=y1.device)
        err = torch.empty((B, Ks), dtype=torch.float64, device=y1.device)
        y0s, y0p = self._y0_args(y0, B)
        self._check(self.lib.dgadj_tdg_adjoint(self._h, B, Ks, NPP, nq, int(self.linear), y0s, y0p,
                                               C.c_void_p(consts.ctypes.data), C.c_void_p(y1.contiguous().data_ptr()),
                                               C.c_void_p(v.data_ptr()), C.c_void_p(err.data_ptr()), self._stream()))
        return nodes, v, err


    # ------------------------------------------------------------------ adj_rec.m
    def adjrec_constants(self, Ns, t1):
        """Per-element block A0 | f1 | R | H | A2 | Ix | np_k | last_{k-1} of the linear branch
        of adj_rec.m (layout: csrc/dgadj_tdg.cu), padded to the largest order of the mesh."""
        Ks = len(t1)
        Ns = self._orders(Ns, Ks)
        els = []
        for k in range(Ks):
            tk = np.asarray(t1[k], dtype=np.float64)
            N = int(Ns[k])
            if N + 1 != len(tk):
                raise _lib.DgadjError(_lib.ERR_INVALID, "adj_rec takes the primal orders (matlab/MAIN.m:35)")
            if N + 1 not in RADAU:
                raise _lib.DgadjError(_lib.ERR_UNSUPPORTED, "utils/Globals1D.m:37-42 tabulates Radau points for N <= 4 only")
            key = ("r", N, tk.tobytes())
            if key not in self._cache:
                g = BaseGalerkin1D(n=N, k=1, domain=(tk[0], tk[-1]), n_gq=1)       # adj_rec.m:20
                x = g.x[:, 0]
                hk = x[0] - x[-1]                                 # :21 (negative)
                M = hk / 2 * g.mass                               # :22
                S = g.mass @ g.d_r                                # :23
                Np = g.n_p
                m = np.zeros((Np, Np)); m[0, 0] = -1.0            # :24
                rad_m = N + 1                                     # :36
                rad_x = tk[0] + (1 + RADAU[rad_m]) * abs(hk) / 2  # :37-38
                x_rec = np.concatenate([rad_x, [tk[-1]]])         # :46
                ge = BaseGalerkin1D(n=rad_m, k=1, domain=(tk[0], tk[-1]), n_gq=1)  # :50 (hk kept)
                xe = ge.x[:, 0]
                me = np.zeros((ge.n_p, ge.n_p)); me[-1, -1] = 1.0                  # :53
                self._cache[key] = dict(
                    Np=Np, x_rec=x_rec, A0=-S.T + m - M, f1=M @ np.ones(Np),       # :25, :31
                    R=_polyfit_matrix(x, Np - 1, rad_x),                           # :42-44
                    H=_polyfit_matrix(x_rec, rad_m, xe),                           # :47-48, :65
                    A2=-(ge.mass @ ge.d_r).T + me + hk / 2 * ge.mass,              # :51-54
                    Ix=_polyfit_matrix(tk, len(tk) - 1, xe))                       # :62-64
            els.append(self._cache[key])
        NP = max(e["Np"] for e in els)
        NA = NP + 1
        blocks = []
        for k, e in enumerate(els):
            parts = [self._pad(e["A0"], NP, NP, identity=True), self._pad(e["f1"], 1, NP), self._pad(e["R"], NP, NP),
                     self._pad(e["H"], NA, NA), self._pad(e["A2"], NA, NA), self._pad(e["Ix"], NA, NP),
                     [float(e["Np"]), float(els[k - 1]["Np"] - 1 if k > 0 else 0)]]
            blocks.append(np.concatenate([np.asarray(p, dtype=np.float64) for p in parts]))
        return np.ascontiguousarray(np.concatenate(blocks)), [e["x_rec"] for e in els], NP

    def adj_rec(self, Ns, Ks, times, y1, t1, y0=1.0):
        """[t, v, err] = adj_rec(Ns, Ks, times)  (matlab/adj_rec.m:1; disabled in the reference,
        matlab/MAIN.m:35): the adjoint solved at the primal order Ns and reconstructed to order
        Ns+1 through the Radau points.  Linear problem (`TimeDG(linear=True)`): adj_rec.m:18-71 on
        the device; returns (t, v[B, Ks, Np_max+1], err[B, Ks]) with t[k] = [Radau points; t_{k+1}].
        Nonlinear (`linear = false`, the setting the file ships with, :11): the reference's branch
        is unfinished -- it returns empty cells and err = 0 (:73-87) -- and so does this."""
        torch = self.torch
        B = y1.shape[0]
        if not self.linear:
            return [None] * Ks, [None] * Ks, torch.zeros((B, Ks), dtype=torch.float64, device=y1.device)
        consts, nodes, NP = self.adjrec_constants(Ns, t1)
        if y1.shape[2] != NP:
            raise _lib.DgadjError(_lib.ERR_INVALID, "y1 must be the [B, Ks, Np_max] array dg_march returned")
        v = torch.empty((B, Ks, NP + 1), dtype=torch.float64, device=y1.device)
        err = torch.empty((B, Ks), dtype=torch.float64, device=y1.device)
        y0s, y0p = self._y0_args(y0, B)
        self._check(self.lib.dgadj_tdg_adjoint_rec(self._h, B, Ks, NP, y0s, y0p, C.c_void_p(consts.ctypes.data),
                                                   C.c_void_p(y1.contiguous().data_ptr()), C.c_void_p(v.data_ptr()),
                                                   C.c_void_p(err.data_ptr()), self._stream()))
        return nodes, v, err


    # ------------------------------------------------------------------ err_contribution.m
    def errcon_weights(self, Ns, t1, NP, nquad=64):
        """cvec[Ks, NP]: err_i = cvec_i . u_i for matlab/err_contribution.m:10-39 -- the integral over
        element i of a(t) (u_h - u_h')(t), a(t) = e^{1-t} - 1 (the dsolve result of :23-25), with u_h
        the degree-Ns(i) polyfit interpolant (:10-14); Gauss quadrature (exact to rounding for a
        polynomial times exp) in place of MATLAB's adaptive `integral` (:39)."""
        Ks = len(t1)
        Ns = self._orders(Ns, Ks)
        xq, wq = np.polynomial.legendre.leggauss(nquad)
        cvec = np.zeros((Ks, NP))
        for i in range(Ks):
            tu = np.asarray(t1[i], dtype=np.float64)
            a, b = tu[0], tu[-1]
            tq = 0.5 * (b - a) * xq + 0.5 * (a + b)
            wa = 0.5 * (b - a) * wq * (np.exp(1.0 - tq) - 1.0)
            eye = np.eye(len(tu))
            for j in range(len(tu)):
                pu = np.polyfit(tu, eye[j], int(Ns[i]))               # :10
                cvec[i, j] = np.sum(wa * (np.polyval(pu, tq) - np.polyval(np.polyder(pu), tq)))   # :13-14,:30-31
        return np.ascontiguousarray(cvec)

    def err_contribution(self, Ks, Ns, uh, t1):
        """[err, res] = err_contribution(Ks, Ns, uh, t1)  (matlab/err_contribution.m:1; unused by the
        reference, MAIN.m:50): per-element error contributions against the exact adjoint of the linear
        model problem.  uh = the [B, Ks, Np_max] array dg_march returned.  Returns err[B, Ks]
        (`res` is an empty cell array in the reference, :4,:47)."""
        torch = self.torch
        B, _, NP = uh.shape
        cvec = self.errcon_weights(Ns, t1, NP)
        err = torch.empty((B, Ks), dtype=torch.float64, device=uh.device)
        self._check(self.lib.dgadj_tdg_err_contribution(self._h, B, Ks, NP, C.c_void_p(cvec.ctypes.data),
                                                        C.c_void_p(uh.contiguous().data_ptr()),
                                                        C.c_void_p(err.data_ptr()), self._stream()))
        return err


def refine(times, Ns, err_mean, n):
    """matlab/MAIN.m:137-141: refine the element with the largest |err| (lowest index on ties,
    quirk C-10) by midpoint insertion; a new order entry n is appended."""
    err_mean = np.asarray(err_mean)
    ref_i = int(np.argmax(np.abs(err_mean)))
    times = np.asarray(times, dtype=np.float64)
    times = np.insert(times, ref_i + 1, 0.5 * (times[ref_i] + times[ref_i + 1]))
    return times, np.append(np.asarray(Ns), n), ref_i
