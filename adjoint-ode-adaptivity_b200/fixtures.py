"""Cross-language fixture I/O: the `*.txt` dumps of utils/Save_to_1D_global_data.m:1-34
(`writematrix(Dr, 'Dr.txt')`, ...), so a MATLAB user can diff this repo's operator set against
the reference's globals, or feed MATLAB-generated operators to the GPU library.

File names, shapes and index conventions are MATLAB's: matrices as comma-separated rows,
index maps 1-based into the column-major flattening of (Np, K) arrays (the in-memory
`BaseGalerkin1D` maps are 0-based / row-major; converted here)."""
from __future__ import annotations

import os

import numpy as np

from .galerkin import BaseGalerkin1D

# utils/Globals1D.m:20-34
_RK4A = [0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
         -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0]
_RK4B = [1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0, 1720146321549.0 / 2090206949498.0,
         3134564353537.0 / 4481467310338.0, 2277821191437.0 / 14882151754819.0]
_RK4C = [0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
         2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0]


def _matlab_ids(flat_rowmajor, Np, K):
    """0-based row-major flat id (i*K + k) -> MATLAB 1-based column-major id (k*Np + i + 1)."""
    f = np.asarray(flat_rowmajor)
    i, k = np.divmod(f, K)
    return k * Np + i + 1


def globals_dict(g: BaseGalerkin1D, dt=None):
    """The variables Save_to_1D_global_data.m writes, keyed by file stem."""
    Np, K = g.n_p, g.k
    vm = _matlab_ids(g.v_map_m, Np, K).T.ravel()        # vmapM(:) of (Nfp, Nfaces, K)
    vp_raw = _matlab_ids(g.v_map_p, Np, K)
    # BaseGalerkin1D leaves unmatched faces at 0; BuildMaps1D.m maps boundary faces to themselves
    vp = np.where(g.v_map_p == 0, _matlab_ids(g.v_map_m, Np, K), vp_raw)
    vp[0, 0] = vm[0]                                    # left boundary face is its own partner
    vp = vp.T.ravel()
    d = dict(Dr=g.d_r, EToE=g.e_to_e + 1, EToF=g.e_to_f + 1, Fmask=g.f_mask + 1, Fscale=g.f_scale, Fx=g.f_x,
             invV=g.inv_v, J=g.j_mat, K=K, LIFT=g.lift, mapB=np.array([1, K * 2]), mapI=1, mapO=K * 2, N=g.n,
             Nfaces=g.n_faces, Nfp=g.n_fp, NODETOL=g.node_tol, Np=Np, nx=g.n_x, r=g.r_lgl[:, None],
             rk4a=_RK4A, rk4b=_RK4B, rk4c=_RK4C, rx=g.r_x, V=g.v, vmapB=np.array([1, K * Np]), vmapI=1, vmapM=vm[:, None],
             vmapO=K * Np, vmapP=vp[:, None], VX=g.v_x[None, :], x=g.x)
    if dt is not None:
        d["dt"] = dt
    return d


def save_globals_txt(g: BaseGalerkin1D, directory, dt=None):
    os.makedirs(directory, exist_ok=True)
    for name, val in globals_dict(g, dt).items():
        a = np.atleast_2d(np.asarray(val))
        fmt = "%d" if np.issubdtype(a.dtype, np.integer) else "%.17g"
        np.savetxt(os.path.join(directory, name + ".txt"), a, delimiter=",", fmt=fmt)


def load_globals_txt(directory):
    """Read a directory written by Save_to_1D_global_data.m (or save_globals_txt)."""
    out = {}
    for fn in sorted(os.listdir(directory)):
        if fn.endswith(".txt"):
            a = np.loadtxt(os.path.join(directory, fn), delimiter=",", ndmin=2)
            out[fn[:-4]] = a if a.size > 1 else a.item()
    return out


def operators_from_globals(d):
    """Operator arrays in the layout dgadj_set_operators expects, from loaded MATLAB globals
    (Mref = inv(V V'))."""
    V = np.asarray(d["V"])
    return dict(Dr=np.ascontiguousarray(d["Dr"]), LIFT=np.ascontiguousarray(d["LIFT"]), V=np.ascontiguousarray(V),
                Mref=np.ascontiguousarray(np.linalg.inv(V @ V.T)), rx=np.ascontiguousarray(d["rx"]),
                Fscale=np.ascontiguousarray(d["Fscale"]), x=np.ascontiguousarray(d["x"]))
