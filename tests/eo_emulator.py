"""NumPy emulation of the arithmetic the CUDA march kernel performs (dgadj_kernels.cuh): the
even/odd basis, the scaled RK residual, the reverse sweep and the checkpointed indicator.
It consumes the operator blocks produced by the C library's host code
(dgadj_host_eo_operators / dgadj_host_eo_prolongation), so the CPU suite checks that host
logic and the kernel's algebra against the oracle without a GPU.  Test infrastructure only."""
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


def eo_prolong(lib, gc, P):
    Np = gc.n_p
    PE, PO = np.zeros(HM * HM), np.zeros(HM * HM)
    viol = C.c_double()
    P = np.ascontiguousarray(P)
    rc = lib.dgadj_host_eo_prolongation(Np, C.c_void_p(P.ctypes.data), C.c_void_p(PE.ctypes.data),
                                        C.c_void_p(PO.ctypes.data), C.byref(viol))
    assert rc == 0
    HEc, HOc, HEf, HOf = (Np + 1) // 2, Np // 2, (Np + 2) // 2, (Np + 1) // 2
    return dict(PE=PE.reshape(HM, HM)[:HEf, :HEc], PO=PO.reshape(HM, HM)[:HOf, :HOc], viol=viol.value)


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


class Level:
    """per-level coefficients {m, q0, q1} (kernel: sm_coef) + blocks"""

    def __init__(self, lib, g, a, dt, alpha, periodic):
        self.b = eo_blocks(lib, g)
        rx = g.r_x[0, :]
        sg = np.sign(a)
        e0 = 0.5 * (-1.0 - (1.0 - alpha) * sg)
        e1 = 0.5 * (1.0 - (1.0 - alpha) * sg)
        self.m = -a * rx * dt
        self.q0 = -g.f_scale[0, :] * e0 / rx
        self.q1 = -g.f_scale[1, :] * e1 / rx
        if not periodic:
            self.q1 = self.q1.copy()
            self.q1[-1] = 0.0
        self.periodic = periodic


def fwd_step(L, ze, zo, rk, uin_fn):
    """ze/zo: (HE,K)/(HO,K) of one trajectory; returns updated copies (kernel: fwd_step)."""
    rka, rkb, rkc = rk
    ze, zo = ze.copy(), zo.copy()
    re, ro = np.zeros_like(ze), np.zeros_like(zo)
    for s in range(len(rka)):
        uF = 0.5 * (ze[0] + zo[0])
        uB = 0.5 * (ze[0] - zo[0])
        uL, uR = np.roll(uB, 1), np.roll(uF, -1)
        if not L.periodic:
            uL[0] = uin_fn(s)
            uR[-1] = uB[-1]
        g0 = (uF - uL) * L.q0
        g1 = (uB - uR) * L.q1
        ge, go = g0 + g1, g0 - g1
        re = rka[s] * re + L.b["DE"] @ zo + L.b["LS"][:, None] * ge
        ro = rka[s] * ro + L.b["DO"] @ ze + L.b["LA"][:, None] * go
        bm = rkb[s] * L.m
        ze = ze + bm * re
        zo = zo + bm * ro
    return ze, zo


def adj_step(L, me, mo, we, wo, rk):
    rka, rkb, _ = rk
    me, mo, we, wo = me.copy(), mo.copy(), we.copy(), wo.copy()
    for s in range(len(rka) - 1, -1, -1):
        bm = rkb[s] * L.m
        we = we + bm * me
        wo = wo + bm * mo
        Ge = L.b["LS"] @ we
        Go = L.b["LA"] @ wo
        gam0 = (Ge + Go) * L.q0
        gam1 = (Ge - Go) * L.q1
        gam1L, gam0R = np.roll(gam1, 1), np.roll(gam0, -1)
        if not L.periodic:
            gam1L[0] = 0.0
            gam0R[-1] = 0.0
        a0, aN = gam0 - gam1L, gam1 - gam0R
        me[0] = me[0] + 0.5 * (a0 + aN)
        mo[0] = mo[0] + 0.5 * (a0 - aN)
        me = me + L.b["DO"].T @ wo
        mo = mo + L.b["DE"].T @ we
        we = rka[s] * we
        wo = rka[s] * wo
    return me, mo, we, wo


def fused(lib, gc, gf, P, u0, a, dt, S, alpha, periodic, rk, jw_c, jw_f, inflow_fn=None, t0=0.0):
    """One trajectory through the fused kernel's algorithm.  Returns dict(uT, J, eta, lam0)."""
    Lc = Level(lib, gc, a, dt, alpha, periodic)
    Lf = Level(lib, gf, a, dt, alpha, periodic)
    pr = eo_prolong(lib, gc, P)
    ze, zo = to_eo(u0)
    fe, fo = pr["PE"] @ ze, pr["PO"] @ zo
    ckpt = []
    time = t0
    for n in range(S):
        uin = (lambda s: inflow_fn(time + rk[2][s] * dt)) if inflow_fn else (lambda s: 0.0)
        se, so = fwd_step(Lf, fe, fo, rk, uin)          # sigma = Phi_f(P u^n)
        ze, zo = fwd_step(Lc, ze, zo, rk, uin)
        time = time + dt
        fe, fo = pr["PE"] @ ze, pr["PO"] @ zo
        ckpt.append((fe - se, fo - so))
    uT = from_eo(ze, zo, 0.5)
    J = np.sum(jw_c * uT)
    me, mo = to_eo(jw_f)
    me, mo = me.copy(), mo.copy()
    h = gf.n_p // 2
    me[:h] *= 0.5
    mo[:h] *= 0.5
    we, wo = np.zeros_like(me), np.zeros_like(mo)
    eta = np.zeros(gc.k)
    for n in range(S - 1, -1, -1):
        re_, ro_ = ckpt[n]
        eta += np.sum(me * re_, axis=0) + np.sum(mo * ro_, axis=0)
        me, mo, we, wo = adj_step(Lf, me, mo, we, wo, rk)
    lam0 = from_eo(me, mo, 1.0)
    return dict(uT=uT, J=J, eta=eta, lam0=lam0, viol=max(Lc.b["viol"], Lf.b["viol"], pr["viol"]))
