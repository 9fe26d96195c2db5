"""Oracle: the advection DG march, its discrete adjoint and the indicator with PER-ELEMENT ORDERS (hp;
SURVEY section 8(f)3: Ns(k) of matlab/MAIN.m:21,141 applied to the DG-in-space march).  TEST INFRASTRUCTURE,
see oracle/__init__.py.  PARITY UNPINNED: the reference has no hp PDE code; this is the scheme of
oracle/advec.py (utils/AdvecRHS1D.m:8-19, the LSERK4 loop of utils/One_code.mlx) with each
element carrying the StartUp1D operators of its OWN order -- written with ragged per-element arrays and a
dense global matrix, i.e. independently of the padded modal formulation the device uses.

Boundary data: periodic, or inflow with uin = 0 (time independent, so one matrix L gives the right-hand side).
"""
from __future__ import annotations

import numpy as np

from . import operators as ops


class HpSpace:
    """Mesh v_x with element orders N_k: per-order reference operators, offsets into the global vector."""

    def __init__(self, orders, v_x):
        self.N = [int(n) for n in orders]
        self.K = len(self.N)
        self.v_x = np.asarray(v_x, dtype=float)
        self.ref = {}
        for n in set(self.N):
            r = ops.JacobiGL(0, 0, n)
            V = ops.Vandermonde1D(n, r)
            self.ref[n] = dict(r=r, V=V, invV=np.linalg.inv(V), Dr=ops.Dmatrix1D(n, r, V), LIFT=ops.Lift1D(n + 1, 2, 1, V),
                               M=ops.mass_matrix(V))
        self.n = [n + 1 for n in self.N]
        self.off = np.concatenate(([0], np.cumsum(self.n)))
        self.h = np.diff(self.v_x)
        self.x = [self.v_x[k] + 0.5 * (self.ref[self.N[k]]["r"] + 1.0) * self.h[k] for k in range(self.K)]

    def sl(self, k):
        return slice(self.off[k], self.off[k + 1])

    def weights(self, psi=None):
        """Weights of J = int psi u dx: (M_k 1)_i J_k psi(x_i)."""
        w = np.zeros(self.off[-1])
        for k in range(self.K):
            wk = (self.ref[self.N[k]]["M"] @ np.ones(self.n[k])) * (self.h[k] / 2.0)
            w[self.sl(k)] = wk if psi is None else wk * psi(self.x[k])
        return w

    def rhs_matrix(self, a, alpha, periodic):
        """L with rhs = L u:  -a rx Dr u + LIFT (Fscale o du),  du = (u^- - u^+) o (a nx - (1-alpha)|a nx|)/2
        (utils/AdvecRHS1D.m:10-19); inflow value 0 at the left end and du(mapO) = 0 when not periodic."""
        nt = self.off[-1]
        L = np.zeros((nt, nt))
        c0 = (a * -1.0 - (1.0 - alpha) * abs(a)) / 2.0
        c1 = (a * 1.0 - (1.0 - alpha) * abs(a)) / 2.0
        for k in range(self.K):
            R = self.ref[self.N[k]]
            rx = 2.0 / self.h[k]
            s = self.sl(k)
            L[s, s] += -a * rx * R["Dr"]
            lift0, lift1 = R["LIFT"][:, 0] * rx * c0, R["LIFT"][:, 1] * rx * c1
            first, last = self.off[k], self.off[k + 1] - 1
            # left face: du = (u_k[0] - u_{k-1}[last]) c0
            L[s, first] += lift0
            if k > 0 or periodic:
                kl = (k - 1) % self.K
                L[s, self.off[kl + 1] - 1] -= lift0
            # right face: du = (u_k[last] - u_{k+1}[0]) c1   (zero at the outflow end)
            if k < self.K - 1 or periodic:
                kr = (k + 1) % self.K
                L[s, last] += lift1
                L[s, self.off[kr]] -= lift1
        return L


def prolongation(c: HpSpace, f: HpSpace):
    """Block-diagonal nodal prolongation coarse -> enriched (each element one order up)."""
    P = np.zeros((f.off[-1], c.off[-1]))
    for k in range(c.K):
        Rc, Rf = c.ref[c.N[k]], f.ref[f.N[k]]
        P[f.sl(k), c.sl(k)] = Rf["V"][:, :c.n[k]] @ Rc["invV"]
    return P


def step(u, L, dt):
    """One LSERK4 step (utils/Globals1D.m:20-34 coefficients; the residual starts at 0: rk4a(1) = 0)."""
    res = np.zeros_like(u)
    for s in range(5):
        res = ops.rk4a[s] * res + dt * (L @ u)
        u = u + ops.rk4b[s] * res
    return u


def step_T(lam, L, dt):
    """Transpose of `step`."""
    lk = np.zeros_like(lam)
    for s in range(4, -1, -1):
        lk = lk + ops.rk4b[s] * lam
        lam = lam + dt * (L.T @ lk)
        lk = ops.rk4a[s] * lk
    return lam


def fwd_adj_indicator(u0, orders, v_x, a, dt, nsteps, alpha=0.0, periodic=True, psi=None):
    """One trajectory.  u0: global coarse vector (element k holds orders[k]+1 nodal values at ITS LGL nodes).
    Returns dict(uT, J, lam0 (enriched, global), eta[K], eta_scale[K], spaces=(c, f))."""
    c = HpSpace(orders, v_x)
    f = HpSpace([n + 1 for n in orders], v_x)
    Lc, Lf = c.rhs_matrix(a, alpha, periodic), f.rhs_matrix(a, alpha, periodic)
    P = prolongation(c, f)
    hist = [np.array(u0, dtype=float)]
    for _ in range(nsteps):
        hist.append(step(hist[-1], Lc, dt))
    J = float(c.weights(psi) @ hist[-1])
    lam = f.weights(psi).copy()
    eta, scale = np.zeros(c.K), np.zeros(c.K)
    for n in range(nsteps - 1, -1, -1):
        pu1 = P @ hist[n + 1]
        uf = step(P @ hist[n], Lf, dt)
        for k in range(c.K):
            s = f.sl(k)
            eta[k] += lam[s] @ (pu1[s] - uf[s])
            scale[k] += np.abs(lam[s]) @ (np.abs(pu1[s]) + np.abs(uf[s]))
        lam = step_T(lam, Lf, dt)
    return dict(uT=hist[-1], J=J, lam0=lam, eta=eta, eta_scale=scale + 1e-300, spaces=(c, f))


def pad(space: HpSpace, u, Nmax):
    """Global ragged vector -> (Nmax+1, K): every element's polynomial evaluated at the LGL nodes of order Nmax."""
    r = ops.JacobiGL(0, 0, Nmax)
    out = np.zeros((Nmax + 1, space.K))
    for k in range(space.K):
        R = space.ref[space.N[k]]
        out[:, k] = ops.Vandermonde1D(space.N[k], r) @ (R["invV"] @ u[space.sl(k)])
    return out


def unpad(space: HpSpace, upad, Nmax):
    """(Nmax+1, K) values of polynomials of the elements' own orders -> the ragged global vector."""
    r = ops.JacobiGL(0, 0, Nmax)
    Vmax = ops.Vandermonde1D(Nmax, r)
    invVmax = np.linalg.inv(Vmax)
    out = np.zeros(space.off[-1])
    for k in range(space.K):
        R = space.ref[space.N[k]]
        out[space.sl(k)] = R["V"] @ (invVmax @ upad[:, k])[:space.n[k]]
    return out


def unpad_covector(space: HpSpace, lpad, Nmax):
    """A covector on the padded nodal values (d J / d u_pad, zero outside the elements' spaces) -> d J / d u_k:
    lam_k = E_k^T lam_pad with E_k = V_max(:, 1:n_k) inv(V_k), the evaluation map `pad` applies."""
    r = ops.JacobiGL(0, 0, Nmax)
    out = np.zeros(space.off[-1])
    for k in range(space.K):
        R = space.ref[space.N[k]]
        E = ops.Vandermonde1D(space.N[k], r) @ R["invV"]
        out[space.sl(k)] = E.T @ lpad[:, k]
    return out
