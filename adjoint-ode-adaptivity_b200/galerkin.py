"""Host-side mirror of the reference's `BaseGalerkin1D` (python/galerkin.py:14-263): same
method names, attribute names and (Np, K) row-major array shapes, NumPy fp64 instead of JAX.

The reference class is an unfinished JAX port that cannot run; the
arithmetic here follows what it is a port *of* -- the MATLAB toolkit in utils/ (StartUp1D.m
chain) and matlab/fem_setup.m -- while keeping galerkin.py's interface.  This is the
once-per-mesh setup; nothing here is on the device hot path, it produces the operator set
`dgadj_set_operators` / `dgadj_set_enriched` upload.

Differences from the reference class that are deliberate:
  * `__init__` accepts the class attributes as keyword overrides (n, k, domain, n_gq, v_x) so
    several meshes can live in one process; with no arguments it behaves like the reference
    (class-level defaults n=1, k=2, domain=[0,1], n_gq=2 -- galerkin.py:18-24).
  * index maps are 0-based into the row-major flattening of (Np, K) arrays.
"""
from __future__ import annotations

import math
from typing import Iterable, Optional

import numpy as np


def gamma(z):
    """python/galerkin.py:10-11."""
    return math.gamma(z)


def _jacobi_table(x, a, b, n):
    """Orthonormal Jacobi polynomials P_0..P_n at the points x, shape (n+1, len(x)).
    Three-term recurrence of utils/JacobiP.m:15-34, all orders in one sweep."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64)).ravel()
    tab = np.empty((n + 1, x.size))
    g0 = 2.0 ** (a + b + 1) / (a + b + 1) * gamma(a + 1) * gamma(b + 1) / gamma(a + b + 1)
    tab[0] = 1.0 / math.sqrt(g0)
    if n >= 1:
        g1 = (a + 1) * (b + 1) / (a + b + 3) * g0
        tab[1] = ((a + b + 2) * x / 2 + (a - b) / 2) / math.sqrt(g1)
    a_old = 2.0 / (2 + a + b) * math.sqrt((a + 1) * (b + 1) / (a + b + 3))
    for m in range(1, n):
        h1 = 2 * m + a + b
        a_new = 2.0 / (h1 + 2) * math.sqrt((m + 1) * (m + 1 + a + b) * (m + 1 + a) * (m + 1 + b)
                                           / (h1 + 1) / (h1 + 3))
        b_new = -(a * a - b * b) / h1 / (h1 + 2)
        tab[m + 1] = (-a_old * tab[m - 1] + (x - b_new) * tab[m]) / a_new
        a_old = a_new
    return tab


class BaseGalerkin1D:
    """Structures shared by all Galerkin methods (python/galerkin.py:14): reference-element
    operators, mesh, geometric factors, connectivity and the quadrature-point basis."""

    n: int = 1
    k: int = 2
    domain: Iterable[float] = np.array([0.0, 1.0])
    n_gq: int = 2
    node_tol: float = 1e-10
    n_fp = 1
    n_faces = 2

    # ------------------------------------------------------------------ reference element
    def jacobiGQ(self, a, b, n):
        """Gauss-Jacobi nodes / weights (galerkin.py:26-45, utils/JacobiGQ.m:8-22)."""
        if n == 0:
            return np.array([-(a - b) / (a + b + 2.0)]), np.array([2.0])
        idx = np.arange(n + 1, dtype=np.float64)
        h1 = 2.0 * idx + a + b
        with np.errstate(divide="ignore", invalid="ignore"):
            diag = -0.5 * (a * a - b * b) / (h1 + 2.0) / h1
        m = idx[1:]
        off = 2.0 / (h1[:-1] + 2.0) * np.sqrt(m * (m + a + b) * (m + a) * (m + b) / (h1[:-1] + 1.0) / (h1[:-1] + 3.0))
        if a + b < 10 * np.finfo(np.float64).eps:
            diag[0] = 0.0
        j_mat = np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)
        d, v = np.linalg.eigh(j_mat)
        w = v[0, :] ** 2 * 2.0 ** (a + b + 1) / (a + b + 1) * gamma(a + 1) * gamma(b + 1) / gamma(a + b + 1)
        return d, w

    def jacobiGL(self, a, b, n):
        """Gauss-Lobatto nodes (galerkin.py:47-52, utils/JacobiGL.m:8-12)."""
        if n == 1:
            return np.array([-1.0, 1.0])
        x_int, _ = self.jacobiGQ(a + 1, b + 1, n - 2)
        return np.concatenate(([-1.0], x_int, [1.0]))

    def jacobiP(self, x, a, b, n):
        """Orthonormal Jacobi polynomial of order n at x (galerkin.py:54-91)."""
        return _jacobi_table(x, a, b, n)[n]

    def vandermonde1D(self, n, r):
        """V[i, j] = P_j(r_i) (galerkin.py:93-97)."""
        return _jacobi_table(r, 0, 0, n).T.copy()

    def gradJacobiP(self, r, a, b, n):
        """galerkin.py:99-102."""
        r = np.atleast_1d(np.asarray(r, dtype=np.float64)).ravel()
        if n == 0:
            return np.zeros(r.size)
        return math.sqrt(n * (n + a + b + 1)) * _jacobi_table(r, a + 1, b + 1, n - 1)[n - 1]

    def gradVandermonde1D(self, n, r):
        """galerkin.py:104-108."""
        r = np.atleast_1d(np.asarray(r, dtype=np.float64)).ravel()
        out = np.zeros((r.size, n + 1))
        if n >= 1:
            tab = _jacobi_table(r, 1, 1, n - 1)
            orders = np.arange(1, n + 1, dtype=np.float64)
            out[:, 1:] = (np.sqrt(orders * (orders + 1.0))[:, None] * tab).T
        return out

    def dMatrix1D(self, n, r, v):
        """Dr = Vr V^-1 (galerkin.py:110-112).  `n` is the polynomial order."""
        v_r = self.gradVandermonde1D(n, r)
        return np.linalg.solve(v.T, v_r.T).T

    def lift1D(self, n_p, n_faces, n_fp=None, v=None):
        """LIFT = V (V^T E) (galerkin.py:114-117)."""
        v = self.v if v is None else v
        e_mat = np.zeros((n_p, n_faces * (self.n_fp if n_fp is None else n_fp)))
        e_mat[0, 0] = 1.0
        e_mat[n_p - 1, 1] = 1.0
        return v @ (v.T @ e_mat)

    def geometricFactors1D(self, x, d_r):
        """Returns (x_r, r_x) in the order of galerkin.py:119-122 (x_r is the Jacobian)."""
        x_r = d_r @ x
        return x_r, 1.0 / x_r

    def normals1D(self):
        """galerkin.py:124-125: outward normals, shape (2, K)."""
        return np.stack((-np.ones(self.k), np.ones(self.k)))

    def connect1D(self, e_to_v):
        """Element-to-element / element-to-face tables (galerkin.py:127-157).  In 1-D the
        face (k, 1) meets face (k+1, 0) wherever they share a vertex; boundary faces keep
        the self reference."""
        k = e_to_v.shape[0]
        e_to_e = np.repeat(np.arange(k)[:, None], self.n_faces, axis=1)
        e_to_f = np.repeat(np.arange(self.n_faces)[None, :], k, axis=0)
        # faces sorted by vertex id: two entries with the same vertex are neighbours
        verts = e_to_v.reshape(-1)
        order = np.argsort(verts, kind="stable")
        same = np.nonzero(verts[order][1:] == verts[order][:-1])[0]
        f1, f2 = order[same], order[same + 1]
        e1, l1 = np.divmod(f1, self.n_faces)
        e2, l2 = np.divmod(f2, self.n_faces)
        e_to_e[e1, l1], e_to_f[e1, l1] = e2, l2
        e_to_e[e2, l2], e_to_f[e2, l2] = e1, l1
        return e_to_e, e_to_f

    def buildMaps1D(self):
        """Face-node index maps into the row-major flattening of (Np, K) arrays
        (galerkin.py:159-196).  v_map_m / v_map_p have shape (n_faces, K)."""
        node_ids = np.arange(self.k * self.n_p).reshape(self.n_p, self.k)
        v_map_m = node_ids[self.f_mask.ravel(), :]                      # (2, K)
        k2, f2 = self.e_to_e.T, self.e_to_f.T                           # (2, K)
        cand = v_map_m[f2, k2]
        xf = self.x.ravel()
        close = (xf[v_map_m] - xf[cand]) ** 2 < self.node_tol
        v_map_p = np.where(close, cand, 0)
        flat_m, flat_p = v_map_m.T.ravel(), v_map_p.T.ravel()           # face-major per element
        map_b = np.nonzero(flat_p == flat_m)[0]
        v_map_b = flat_m[map_b]
        self.map_i = 0
        self.map_o = self.k * self.n_faces - 1
        self.v_map_i = 0
        self.v_map_o = self.k * self.n_p - 1
        return v_map_m, v_map_p, v_map_b, map_b

    # ------------------------------------------------------------------ setup
    def startUp1D(self):
        """galerkin.py:199-237 / utils/StartUp1D.m:5-39."""
        self.n_p = self.n + 1
        self.r = self.jacobiGL(0, 0, self.n)
        self.v = self.vandermonde1D(self.n, self.r)
        self.inv_v = np.linalg.inv(self.v)
        self.d_r = self.dMatrix1D(self.n, self.r, self.v)
        self.lift = self.lift1D(self.n_p, self.n_faces)
        v_a = self.e_to_v[:, 0]
        v_b = self.e_to_v[:, 1]
        self.x = np.ones((self.n_p, 1)) * self.v_x[v_a][None, :] \
            + 0.5 * (self.r[:, None] + 1.0) * (self.v_x[v_b] - self.v_x[v_a])[None, :]
        self.j_mat, self.r_x = self.geometricFactors1D(self.x, self.d_r)
        self.f_mask = np.stack((np.nonzero(np.abs(self.r + 1) < self.node_tol)[0],
                                np.nonzero(np.abs(self.r - 1) < self.node_tol)[0]), axis=1)  # (1, 2)
        self.f_x = self.x[self.f_mask.ravel(), :]
        self.n_x = self.normals1D()
        self.f_scale = 1.0 / self.j_mat[self.f_mask.ravel(), :]
        self.e_to_e, self.e_to_f = self.connect1D(self.e_to_v)
        self.v_map_m, self.v_map_p, self.v_map_b, self.map_b = self.buildMaps1D()

    def __init__(self, n: Optional[int] = None, k: Optional[int] = None, domain=None,
                 n_gq: Optional[int] = None, v_x=None) -> None:
        if n is not None:
            self.n = int(n)
        if domain is not None:
            self.domain = np.asarray(domain, dtype=np.float64)
        if n_gq is not None:
            self.n_gq = int(n_gq)
        if v_x is not None:                      # arbitrary (refined, non-uniform) vertex list
            self.v_x = np.asarray(v_x, dtype=np.float64)
            self.k = self.v_x.size - 1
            self.domain = np.array([self.v_x[0], self.v_x[-1]])
        else:
            if k is not None:
                self.k = int(k)
            nv = self.k + 1
            dom = np.asarray(self.domain, dtype=np.float64)
            # utils/MeshGen1D.m:8: VX(i) = (xmax-xmin)*(i-1)/(Nv-1) + xmin
            self.v_x = (dom[1] - dom[0]) * np.arange(nv) / (nv - 1) + dom[0]
        self.e_to_v = np.stack((np.arange(self.k), np.arange(1, self.k + 1)), axis=1)
        self.startUp1D()
        self.r_lgl = self.r
        # quadrature rule + nodal basis at the quadrature points (galerkin.py:249-263,
        # matlab/fem_setup.m:27-39); as in the reference `r` is overwritten (quirk C-7)
        self.r, self.w = self.jacobiGQ(0, 0, self.n_gq)
        self.n_r = self.r.shape[0]
        self.phi = _jacobi_table(self.r, 0, 0, self.n).T @ self.inv_v   # Phi[q, i] = l_i(r_q)

    # ------------------------------------------------------------------ derived operators
    @property
    def mass(self):
        """Reference mass matrix M = (V V^T)^-1 (matlab/dg_march.m:15)."""
        return np.linalg.inv(self.v @ self.v.T)

    def prolongation_to(self, fine: "BaseGalerkin1D"):
        """Nodal prolongation to a higher-order space on the same mesh:
        P = V_fine(:, 0:Np) V^-1."""
        v_low = _jacobi_table(fine.r_lgl, 0, 0, self.n).T
        return v_low @ self.inv_v

    def quad_weights(self):
        """Nodal weights of int u dx: (M_k 1)_i = J[i,k] (Mref 1)_i, shape (Np, K)."""
        return (self.mass @ np.ones(self.n_p))[:, None] * self.j_mat
