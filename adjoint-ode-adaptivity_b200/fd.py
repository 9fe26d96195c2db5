"""Finite-difference path of the reference (python/Main_finite_difference.py:16-94, 263-343) on
the GPU, batched over initial conditions on a shared time mesh.

The reference passes Python callables (`updateRule`, `getJF`, `getK`); a device kernel cannot
run those, so the two ODEs and three output functionals its `__main__` block defines are
selected by name (`ode` in {"sin", "linear"}, `functional` in {"int_u", "u_N", "int_u2"}).
Function names and array meanings follow the reference: `forwardSolve` -> u[n+1],
`adjSolve` -> v on the `ref_factor`-refined mesh, `errEst` -> signed fine-mesh indicator.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def refineAll(dt_n, ref_factor):
    """Main_finite_difference.py:16-21 (host helper: the mesh is shared by the batch)."""
    dt_n = np.asarray(dt_n, dtype=np.float64)
    return np.repeat(dt_n / ref_factor, ref_factor), dt_n.size * ref_factor


class FDAdjoint:
    """Device handle for the FD path.  `u0` is a torch CUDA tensor [B] (float64)."""

    def __init__(self, ode="sin", functional="int_u2", ref_factor=4, device=0):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.ode, self.functional, self.ref_factor, self.device = ode, functional, int(ref_factor), device
        # the march handle carries device / error state; its DG fields are unused here
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

    def solve(self, u0, dt_n, want=("u", "v", "err_fine", "err_steps", "ref_idx")):
        """forwardSolve -> adjSolve -> errEst -> window sums -> argmax in one launch.
        Returns a dict with the requested arrays: u[B, n+1], v[B, n*rf+1], err_fine[B, n*rf+1],
        err_steps[B, n], ref_idx[B] (int32, 0-based element to refine)."""
        torch = self.torch
        if not (isinstance(u0, torch.Tensor) and u0.is_cuda and u0.dtype == torch.float64):
            raise TypeError("u0 must be a float64 CUDA tensor of shape [B]")
        u0 = u0.contiguous().view(-1)
        dt_n = np.ascontiguousarray(dt_n, dtype=np.float64)
        B, n, nf = u0.numel(), dt_n.size, dt_n.size * self.ref_factor
        kw = dict(dtype=torch.float64, device=u0.device)
        out = {}
        if "u" in want:
            out["u"] = torch.empty((B, n + 1), **kw)
        if "v" in want:
            out["v"] = torch.empty((B, nf + 1), **kw)
        if "err_fine" in want:
            out["err_fine"] = torch.empty((B, nf + 1), **kw)
        if "err_steps" in want:
            out["err_steps"] = torch.empty((B, n), **kw)
        if "ref_idx" in want:
            out["ref_idx"] = torch.empty(B, dtype=torch.int32, device=u0.device)
        p = lambda k: C.c_void_p(out[k].data_ptr()) if k in out else C.c_void_p(0)
        rc = self.lib.dgadj_fd_awr(self._h, B, n, self.ref_factor, _lib.FD_ODE[self.ode],
                                   _lib.FD_FUNCTIONAL[self.functional], C.c_void_p(dt_n.ctypes.data),
                                   C.c_void_p(u0.data_ptr()), p("u"), p("v"), p("err_fine"), p("err_steps"),
                                   p("ref_idx"), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != _lib.OK:
            raise _lib.DgadjError(rc, self.lib.dgadj_last_error(self._h).decode())
        return out

    # reference-named views of the single fused launch -------------------------------------
    def forwardSolve(self, dt_n, u0):
        """python/Main_finite_difference.py:34-51."""
        return self.solve(u0, dt_n, want=("u",))["u"]

    def adjSolve(self, dt_n, u0):
        """python/Main_finite_difference.py:54-76 (needs the primal: marches it first)."""
        return self.solve(u0, dt_n, want=("v",))["v"]

    def errEst(self, dt_n, u0):
        """python/Main_finite_difference.py:79-94."""
        return self.solve(u0, dt_n, want=("err_fine",))["err_fine"]


def refine_mesh(times, ref_idx):
    """Insert the midpoint of element `ref_idx` (python/Main_finite_difference.py:336-341;
    matlab/MAIN.m:138-141)."""
    times = np.asarray(times, dtype=np.float64)
    mid = 0.5 * (times[ref_idx] + times[ref_idx + 1])
    return np.insert(times, ref_idx + 1, mid)
