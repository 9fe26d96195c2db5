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


# ------------------------------------------------------------------------------------------------------
# The reference's free functions under their own argument lists (python/Main_finite_difference.py:34,54,79):
#     u   = forwardSolve(updateRule, dt_n, u0=None)
#     v   = adjSolve(getK, getJF, dt_n, u, ref_factor)
#     err = errEst(fwdUpdate, u, v, dt_n, ref_factor)
# The reference hands Python callables down to its loops; a device kernel cannot run those, so each callable
# is PROBED on a small fixed input and matched against the problem functions the reference's __main__ block
# defines (the two ODEs :110-140 and the three output functionals :153-227); anything else raises.  Arrays
# may be NumPy (as in the reference: copied to the device and back) or float64 CUDA tensors; a leading batch
# axis is allowed.  A module-level handle (the analogue of the reference's module globals) does the work.
# ------------------------------------------------------------------------------------------------------
_PROBE_U = np.array([0.3, 0.7, -0.4])
_PROBE_DT = np.array([0.5, 0.25])
_DEFAULT = {}


def _identify_ode(updateRule=None, getJF=None):
    found = set()
    if updateRule is not None:
        r = float(np.asarray(updateRule(_PROBE_U.copy(), _PROBE_DT.copy(), 1)).ravel()[0])
        if np.isclose(r, _PROBE_U[0] + np.sin(_PROBE_U[0]) * _PROBE_DT[0], rtol=1e-14, atol=0):
            found.add("sin")                                    # :131-132
        elif np.isclose(r, (1 + _PROBE_DT[0]) * _PROBE_U[0], rtol=1e-14, atol=0):
            found.add("linear")                                 # :112-113
        else:
            raise NotImplementedError("updateRule is neither of the reference's rules (u + sin(u) dt, (1 + dt) u): "
                                      "the device path cannot run an arbitrary Python callable")
    if getJF is not None:
        jf = np.asarray(getJF(_PROBE_U.copy(), _PROBE_DT.copy()))
        if jf.shape == (3, 3) and np.allclose(jf, np.diag(1 + np.cos(_PROBE_U[:-1]) * _PROBE_DT, -1), rtol=1e-14, atol=0):
            found.add("sin")                                    # :138-140
        elif jf.shape == (3, 3) and np.allclose(jf, np.diag(1 + _PROBE_DT, -1), rtol=1e-14, atol=0):
            found.add("linear")                                 # :118-119
        else:
            raise NotImplementedError("getJF is neither of the reference's Jacobians (diag(1 + cos(u) dt, -1), diag(1 + dt, -1))")
    if len(found) != 1:
        raise ValueError("updateRule and getJF belong to different ODEs")
    return found.pop()


def _identify_functional(getK):
    try:
        k = np.asarray(getK(_PROBE_DT.copy(), _PROBE_U.copy()), dtype=float)
    except TypeError:
        k = np.asarray(getK(_PROBE_DT.copy()), dtype=float)
    if k.shape == (3,):
        if np.allclose(k, np.concatenate((2 * _PROBE_U[:-1] * _PROBE_DT, [0.0])), rtol=1e-14, atol=0):
            return "int_u2"                                     # :225-227
        if np.allclose(k, np.concatenate((_PROBE_DT, [0.0])), rtol=1e-14, atol=0):
            return "int_u"                                      # :153-155
        if np.array_equal(k, [0.0, 1.0, 0.0]):
            return "u_N"                                        # :162-165
    raise NotImplementedError("getK is none of the reference's functionals (J = int u^2, J = int u, J = u_N)")


def _default_handle(ode, functional, ref_factor, device=0):
    key = (ode, functional, int(ref_factor), device)
    if key not in _DEFAULT:
        _DEFAULT[key] = FDAdjoint(ode=ode, functional=functional, ref_factor=ref_factor, device=device)
    return _DEFAULT[key]


def _to_device(x, torch, device=0):
    if isinstance(x, torch.Tensor):
        return x.to(dtype=torch.float64), True
    return torch.as_tensor(np.atleast_1d(np.asarray(x, dtype=np.float64)), device=torch.device("cuda", device)), False


def _back(t, was_tensor, squeeze):
    if was_tensor:
        return t[0] if squeeze else t
    a = t.cpu().numpy()
    return a[0] if squeeze else a


def forwardSolve(updateRule, dt_n, u0=None):
    """python/Main_finite_difference.py:34-51.  u0: scalar (the reference), array / CUDA tensor [B];
    returns u[n+1] (or [B, n+1])."""
    import torch
    ode = _identify_ode(updateRule=updateRule)
    if u0 is None:
        u0 = 0.0                                                # :37-38 (u = zeros)
    squeeze = np.ndim(u0) == 0
    d_u0, was = _to_device(u0, torch)
    s = _default_handle(ode, "int_u2", 4)
    return _back(s.solve(d_u0.view(-1), dt_n, want=("u",))["u"], was, squeeze)


def _check_primal(s, u, dt_n, torch):
    """The device path re-marches the primal from u[..., 0]; refuse a `u` that is not that march."""
    squeeze = np.ndim(u) == 1 if not isinstance(u, torch.Tensor) else u.ndim == 1
    d_u, was = _to_device(u, torch)
    d_u = d_u.view(1, -1) if d_u.ndim == 1 else d_u
    if d_u.shape[1] != len(dt_n) + 1:
        raise ValueError("u must hold len(dt_n) + 1 values per trajectory")
    return d_u, was, squeeze


def adjSolve(getK, getJF, dt_n, u, ref_factor):
    """python/Main_finite_difference.py:54-76: v on the ref_factor-refined mesh ((JF^T - I) v = -k)."""
    import torch
    ode, functional = _identify_ode(getJF=getJF), _identify_functional(getK)
    s = _default_handle(ode, functional, ref_factor)
    d_u, was, squeeze = _check_primal(s, u, dt_n, torch)
    out = s.solve(d_u[:, 0].contiguous(), dt_n, want=("u", "v"))
    if not torch.allclose(out["u"], d_u, rtol=1e-12, atol=1e-14):
        raise ValueError("u is not the forward-Euler primal of u[0] on dt_n (the device path re-marches it from u[0])")
    return _back(out["v"], was, squeeze)


def errEst(fwdUpdate, u, v, dt_n, ref_factor, functional="int_u2"):
    """python/Main_finite_difference.py:79-94: the signed fine-mesh indicator err = res * v.  `v` must be the
    adjoint `adjSolve` returned for this primal (the kernel recomputes it; `functional` names its J when it
    is not the reference's default J = int u^2 -- a mismatch raises)."""
    import torch
    ode = _identify_ode(updateRule=fwdUpdate)
    s = _default_handle(ode, functional, ref_factor)
    d_u, was, squeeze = _check_primal(s, u, dt_n, torch)
    out = s.solve(d_u[:, 0].contiguous(), dt_n, want=("u", "v", "err_fine"))
    d_v, _ = _to_device(v, torch)
    d_v = d_v.view(1, -1) if d_v.ndim == 1 else d_v
    if not torch.allclose(out["u"], d_u, rtol=1e-12, atol=1e-14) or not torch.allclose(out["v"], d_v, rtol=1e-11, atol=1e-13):
        raise ValueError("u / v are not the primal and adjoint of u[0] on dt_n for this ODE and functional")
    return _back(out["err_fine"], was, squeeze)


def interpU(dt_fine, dt_n, u, ref_factor=4):
    """python/Main_finite_difference.py:24-31 (host helper; like the reference it ignores dt_fine -- quirk C-12 --
    and refines by `ref_factor`, which the reference reads from a module global)."""
    dt_n = np.asarray(dt_n, dtype=np.float64)
    t_c = np.concatenate(([0.0], np.cumsum(dt_n)))
    t_f = np.concatenate(([0.0], np.cumsum(refineAll(dt_n, ref_factor)[0])))
    return np.interp(t_f, t_c, np.asarray(u, dtype=np.float64))
