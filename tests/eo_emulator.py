"""The even/odd (symmetric / antisymmetric) operator blocks the Burgers kernels use for Dr f:
helpers around the C library's host code dgadj_host_eo_operators.  Test infrastructure only."""
import ctypes as C

import numpy as np

HM = 5


def eo_blocks(lib, g):
    Np = g.n_p
    DE, DO = np.zeros(HM * HM), np.zeros(HM * HM)
    LS, LA = np.zeros(HM), np.zeros(HM)
    viol = C.c_double()
    p = lambda a: C.c_void_p(a.ctypes.data)
    Dr, LIFT = np.ascontiguousarray(g.d_r), np.ascontiguousarray(g.lift)
    rc = lib.dgadj_host_eo_operators(Np, p(Dr), p(LIFT), p(DE), p(DO), p(LS), p(LA), C.byref(viol))
    assert rc == 0
    HE, HO = (Np + 1) // 2, Np // 2
    return dict(DE=DE.reshape(HM, HM)[:HE, :HO], DO=DO.reshape(HM, HM)[:HO, :HE], LS=LS[:HE], LA=LA[:HO],
                viol=viol.value, HE=HE, HO=HO, Np=Np)


def to_eo(u):
    Np = u.shape[-2]
    h = Np // 2
    e = [u[..., i, :] + u[..., Np - 1 - i, :] for i in range(h)]
    o = [u[..., i, :] - u[..., Np - 1 - i, :] for i in range(h)]
    if Np & 1:
        e.append(u[..., h, :])
    return np.stack(e, axis=-2), np.stack(o, axis=-2)


def from_eo(e, o, half):
    HE, HO = e.shape[-2], o.shape[-2]
    Np = HE + HO
    u = np.zeros(e.shape[:-2] + (Np, e.shape[-1]))
    for i in range(HO):
        u[..., i, :] = half * (e[..., i, :] + o[..., i, :])
        u[..., Np - 1 - i, :] = half * (e[..., i, :] - o[..., i, :])
    if Np & 1:
        u[..., HO, :] = e[..., HO, :]
    return u
