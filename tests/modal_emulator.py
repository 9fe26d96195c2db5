"""NumPy emulation of the arithmetic the CUDA march kernel performs (dgadj_kernels.cuh): the
modal (orthonormal Legendre) basis, the parity-sparse derivative, the stage-scaled RK residual,
the reverse sweep and the checkpointed indicator.  It consumes the operators produced by the C
library's host code (dgadj_host_modal_operators), so the CPU suite checks that host logic and
the kernel's algebra against the oracle without a GPU.  Test infrastructure only."""
import ctypes as C

import numpy as np

MAXNZ = 26


def nz_index(Np, i, j):
    return sum((Np - q) // 2 for q in range(i)) + (j - i - 1) // 2


def modal_ops(lib, g):
    Np = g.n_p
    Dnz, p, iV = np.zeros(MAXNZ), np.zeros(Np), np.zeros((Np, Np))
    viol = C.c_double()
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    Dr, LIFT, V = c(g.d_r), c(g.lift), c(g.v)
    ptr = lambda a: C.c_void_p(a.ctypes.data)
    rc = lib.dgadj_host_modal_operators(Np, ptr(Dr), ptr(LIFT), ptr(V), ptr(Dnz), ptr(p), ptr(iV), C.byref(viol))
    assert rc == 0
    D = np.zeros((Np, Np))
    for i in range(Np):
        for j in range(i + 1, Np, 2):
            D[i, j] = Dnz[nz_index(Np, i, j)]
    return dict(D=D, p=p, V=V, iV=iV, viol=viol.value, Np=Np)


class Level:
    """per-level coefficients {m, q0, q1} (kernel: sm_coef) + operators"""

    def __init__(self, lib, g, a, dt, alpha, periodic):
        self.o = modal_ops(lib, g)
        rx = g.r_x.sum(axis=0) / g.r_x.shape[0]        # the kernel's per-element rx: mean over the nodes
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


def stage_scalings(rka):
    ns = len(rka)
    sig, sga = np.ones(ns), np.ones(ns)
    for s in range(1, ns):
        sig[s] = rka[s] * sig[s - 1]
    for s in range(ns - 2, -1, -1):
        sga[s] = rka[s + 1] * sga[s + 1]
    return sig, sga


def fwd_step(L, z, rk, uin_fn, mask=None):
    """z: (Np, K) modal coefficients of one trajectory (kernel: fwd_step).  mask (Np, K) bool: hp -- the modes of
    each element's own space; the surface term and the update skip the others (kernel: `HP && i >= nm[e]`)."""
    mk = 1.0 if mask is None else mask.astype(float)
    rka, rkb, _ = rk
    o, p = L.o, L.o["p"]
    ev = (np.arange(len(p)) % 2 == 0)
    sig, _ = stage_scalings(rka)
    z = z.copy()
    r = np.zeros_like(z)
    for s in range(len(rka)):
        Se, So = (p[ev, None] * z[ev]).sum(0), (p[~ev, None] * z[~ev]).sum(0)
        uF, uB = Se - So, Se + So
        uL, uR = np.roll(uB, 1), np.roll(uF, -1)
        if not L.periodic:
            uL[0] = uin_fn(s)
            uR[-1] = uB[-1]
        g0, g1 = (uF - uL) * L.q0, (uB - uR) * L.q1
        se, sd = (g1 + g0) / sig[s], (g1 - g0) / sig[s]
        r = r + (o["D"] / sig[s]) @ z + mk * (p[:, None] * np.where(ev[:, None], se, sd))
        z = z + mk * ((rkb[s] * sig[s] * L.m) * r)
    return z


def adj_step(L, mu, rk, mask=None):
    mk = 1.0 if mask is None else mask.astype(float)
    rka, rkb, _ = rk
    o, p = L.o, L.o["p"]
    ev = (np.arange(len(p)) % 2 == 0)
    _, sga = stage_scalings(rka)
    mu = mu.copy()
    w = np.zeros_like(mu)
    for s in range(len(rka) - 1, -1, -1):
        w = w + (rkb[s] / sga[s] * L.m) * mu
        Gse, Gso = (p[ev, None] * w[ev]).sum(0), (p[~ev, None] * w[~ev]).sum(0)
        gam0, gam1 = (Gse - Gso) * (sga[s] * L.q0), (Gse + Gso) * (sga[s] * L.q1)
        gam1L, gam0R = np.roll(gam1, 1), np.roll(gam0, -1)
        if not L.periodic:
            gam1L[0] = 0.0
            gam0R[-1] = 0.0
        a0, aN = gam0 - gam1L, gam1 - gam0R
        mu = mu + mk * ((o["D"] * sga[s]).T @ w) + mk * (p[:, None] * np.where(ev[:, None], aN + a0, aN - a0))
    return mu


def fused(lib, gc, gf, u0, a, dt, S, alpha, periodic, rk, jw_c, jw_f, inflow_fn=None, t0=0.0, nodes_per_element=None):
    """One trajectory through the fused kernel's algorithm.  Returns dict(uT, J, eta, lam0).
    nodes_per_element (K ints <= Np): hp, the HP kernels' mode masks (the enriched space has one mode more)."""
    Lc, Lf = Level(lib, gc, a, dt, alpha, periodic), Level(lib, gf, a, dt, alpha, periodic)
    K = u0.shape[1]
    mc = mf = None
    if nodes_per_element is not None:
        nm = np.asarray(nodes_per_element)
        mc = np.arange(gc.n_p)[:, None] < nm[None, :]
        mf = np.arange(gf.n_p)[:, None] < (nm[None, :] + 1)
    zc = Lc.o["iV"] @ u0
    if mc is not None:
        zc = zc * mc                     # the L2 projection of the input onto each element's own space
    ckpt = []
    time = t0
    for n in range(S):
        uin = (lambda s: inflow_fn(time + rk[2][s] * dt)) if inflow_fn else (lambda s: 0.0)
        sig = fwd_step(Lf, np.vstack([zc, np.zeros((1, K))]), rk, uin, mf)     # sigma = Phi_f(P u^n), P = injection
        zc = fwd_step(Lc, zc, rk, uin, mc)
        time = time + dt
        ckpt.append(np.vstack([zc, np.zeros((1, K))]) - sig)
    uT = Lc.o["V"] @ zc
    J = np.sum((Lc.o["V"].T @ jw_c) * zc)
    mu = Lf.o["V"].T @ jw_f
    if mf is not None:
        mu = mu * mf                     # the functional's covector restricted to the elements' enriched spaces
    eta = np.zeros(K)
    for n in range(S - 1, -1, -1):
        eta += np.sum(mu * ckpt[n], axis=0)
        mu = adj_step(Lf, mu, rk, mf)
    lam0 = Lf.o["iV"].T @ mu
    return dict(uT=uT, J=J, eta=eta, lam0=lam0, viol=max(Lc.o["viol"], Lf.o["viol"]))
