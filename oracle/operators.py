"""Oracle: reference-element + mesh setup (TEST INFRASTRUCTURE, see oracle/__init__.py).

Line-by-line NumPy restatement of the Hesthaven-Warburton 1-D toolkit the reference
ships in `utils/`.  0-based indices, arrays shaped (Np, K) with u[i, k] = node i of
element k; MATLAB's column-major flat numbering is reproduced where it is observable
(vmapM / vmapP / mapI / mapO are returned 1-based, exactly as the mlx prints them).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

NODETOL = 1e-10  # utils/StartUp1D.m:5

# utils/Globals1D.m:20-34 -- low-storage RK (Carpenter-Kennedy) coefficients
rk4a = np.array([
    0.0,
    -567301805773.0 / 1357537059087.0,
    -2404267990393.0 / 2016746695238.0,
    -3550918686646.0 / 2091501179385.0,
    -1275806237668.0 / 842570457699.0,
])
rk4b = np.array([
    1432997174477.0 / 9575080441755.0,
    5161836677717.0 / 13612068292357.0,
    1720146321549.0 / 2090206949498.0,
    3134564353537.0 / 4481467310338.0,
    2277821191437.0 / 14882151754819.0,
])
rk4c = np.array([
    0.0,
    1432997174477.0 / 9575080441755.0,
    2526269341429.0 / 6820363962896.0,
    2006345519317.0 / 3224310063776.0,
    2802321613138.0 / 2924317926251.0,
])


def JacobiGQ(alpha, beta, N):
    """utils/JacobiGQ.m:8-22 -- Gauss-Jacobi nodes/weights via the Jacobi matrix."""
    if N == 0:
        return (np.array([-(alpha - beta) / (alpha + beta + 2.0)]), np.array([2.0]))
    h1 = 2.0 * np.arange(N + 1) + alpha + beta
    with np.errstate(divide="ignore", invalid="ignore"):
        d0 = -0.5 * (alpha**2 - beta**2) / (h1 + 2.0) / h1          # :13
    n = np.arange(1, N + 1, dtype=float)
    d1 = 2.0 / (h1[:N] + 2.0) * np.sqrt(n * (n + alpha + beta) * (n + alpha) * (n + beta)
                                        / (h1[:N] + 1.0) / (h1[:N] + 3.0))  # :14-15
    J = np.diag(d0) + np.diag(d1, 1)
    if alpha + beta < 10 * np.finfo(float).eps:                       # :16
        J[0, 0] = 0.0
    J = J + J.T                                                       # :17
    D, V = np.linalg.eigh(J)                                          # :20 (symmetric eig, ascending)
    x = D
    w = (V[0, :] ** 2) * 2.0 ** (alpha + beta + 1) / (alpha + beta + 1) \
        * math.gamma(alpha + 1) * math.gamma(beta + 1) / math.gamma(alpha + beta + 1)  # :21-22
    return x, w


def JacobiGL(alpha, beta, N):
    """utils/JacobiGL.m:8-12 -- Legendre-Gauss-Lobatto nodes."""
    if N == 1:
        return np.array([-1.0, 1.0])
    xint, _ = JacobiGQ(alpha + 1, beta + 1, N - 2)
    return np.concatenate(([-1.0], xint, [1.0]))


def JacobiP(x, alpha, beta, N):
    """utils/JacobiP.m:12-36 -- orthonormal Jacobi polynomial P_N^(alpha,beta)(x)."""
    xp = np.atleast_1d(np.asarray(x, dtype=float)).ravel()
    PL = np.zeros((N + 1, xp.size))
    gamma0 = 2.0 ** (alpha + beta + 1) / (alpha + beta + 1) * math.gamma(alpha + 1) \
        * math.gamma(beta + 1) / math.gamma(alpha + beta + 1)         # :15-16
    PL[0, :] = 1.0 / math.sqrt(gamma0)                                # :17
    if N == 0:
        return PL[0, :].copy()
    gamma1 = (alpha + 1) * (beta + 1) / (alpha + beta + 3) * gamma0   # :19
    PL[1, :] = ((alpha + beta + 2) * xp / 2 + (alpha - beta) / 2) / math.sqrt(gamma1)  # :20
    if N == 1:
        return PL[1, :].copy()
    aold = 2.0 / (2 + alpha + beta) * math.sqrt((alpha + 1) * (beta + 1) / (alpha + beta + 3))  # :24
    for i in range(1, N):                                             # :27-34
        h1 = 2 * i + alpha + beta
        anew = 2.0 / (h1 + 2) * math.sqrt((i + 1) * (i + 1 + alpha + beta) * (i + 1 + alpha)
                                          * (i + 1 + beta) / (h1 + 1) / (h1 + 3))
        bnew = -(alpha**2 - beta**2) / h1 / (h1 + 2)
        PL[i + 1, :] = 1.0 / anew * (-aold * PL[i - 1, :] + (xp - bnew) * PL[i, :])
        aold = anew
    return PL[N, :].copy()


def Vandermonde1D(N, r):
    """utils/Vandermonde1D.m:6-9 -- V[i, j] = P_j(r_i)."""
    r = np.asarray(r, dtype=float).ravel()
    V = np.zeros((r.size, N + 1))
    for j in range(N + 1):
        V[:, j] = JacobiP(r, 0, 0, j)
    return V


def GradJacobiP(r, alpha, beta, N):
    """utils/GradJacobiP.m:7-12."""
    r = np.asarray(r, dtype=float).ravel()
    if N == 0:
        return np.zeros(r.size)
    return math.sqrt(N * (N + alpha + beta + 1)) * JacobiP(r, alpha + 1, beta + 1, N - 1)


def GradVandermonde1D(N, r):
    """utils/GradVandermonde1D.m:6-11."""
    r = np.asarray(r, dtype=float).ravel()
    DVr = np.zeros((r.size, N + 1))
    for i in range(N + 1):
        DVr[:, i] = GradJacobiP(r, 0, 0, i)
    return DVr


def Dmatrix1D(N, r, V):
    """utils/Dmatrix1D.m:7-8 -- Dr = Vr / V  (right division: solve V' Dr' = Vr')."""
    Vr = GradVandermonde1D(N, r)
    return np.linalg.solve(V.T, Vr.T).T


def Lift1D(Np, Nfaces, Nfp, V):
    """utils/Lift1D.m:7-13 -- LIFT = V (V' E)."""
    Emat = np.zeros((Np, Nfaces * Nfp))
    Emat[0, 0] = 1.0
    Emat[Np - 1, 1] = 1.0
    return V @ (V.T @ Emat)


def GeometricFactors1D(x, Dr):
    """utils/GeometricFactors1D.m:6."""
    J = Dr @ x
    return 1.0 / J, J


def Normals1D(K):
    """utils/Normals1D.m:7-10."""
    nx = np.zeros((2, K))
    nx[0, :] = -1.0
    nx[1, :] = 1.0
    return nx


def MeshGen1D(xmin, xmax, K):
    """utils/MeshGen1D.m:4-14 (EToV returned 0-based)."""
    Nv = K + 1
    VX = np.array([(xmax - xmin) * i / (Nv - 1) + xmin for i in range(Nv)])
    EToV = np.stack([np.arange(K), np.arange(1, K + 1)], axis=1)
    return Nv, VX, K, EToV


def Connect1D(EToV):
    """utils/Connect1D.m:9-40 -- element/face connectivity (returned 0-based).

    The reference builds a sparse face-to-vertex incidence and finds faces sharing a
    vertex; in 1-D that is: face (k,1) <-> face (k+1,0).  Unmatched (boundary) faces
    keep the self-reference defaults of :38-39.
    """
    K = EToV.shape[0]
    Nfaces = 2
    EToE = np.repeat(np.arange(K)[:, None], Nfaces, axis=1)
    EToF = np.repeat(np.arange(Nfaces)[None, :], K, axis=0)
    owner = {}
    for k in range(K):
        for f in range(Nfaces):
            v = int(EToV[k, f])
            if v in owner:
                k2, f2 = owner[v]
                EToE[k, f], EToF[k, f] = k2, f2
                EToE[k2, f2], EToF[k2, f2] = k, f
            else:
                owner[v] = (k, f)
    return EToE, EToF


def BuildMaps1D(Np, K, Fmask, EToE, EToF, x):
    """utils/BuildMaps1D.m:10-43 -- returns MATLAB-numbered (1-based, column-major) maps."""
    nodeids = np.arange(1, K * Np + 1).reshape(K, Np).T       # reshape(1:K*Np, Np, K)
    Nfaces = 2
    vmapM = np.zeros((Nfaces, K), dtype=np.int64)
    vmapP = np.zeros((Nfaces, K), dtype=np.int64)
    for k1 in range(K):
        for f1 in range(Nfaces):
            vmapM[f1, k1] = nodeids[Fmask[f1], k1]
    xf = x.T.ravel()                                          # column-major flat x
    for k1 in range(K):
        for f1 in range(Nfaces):
            k2, f2 = EToE[k1, f1], EToF[k1, f1]
            vidM, vidP = vmapM[f1, k1], vmapM[f2, k2]
            D = (xf[vidM - 1] - xf[vidP - 1]) ** 2
            if D < NODETOL:
                vmapP[f1, k1] = vidP
    vmapP = vmapP.T.ravel()                                   # vmapP(:) of (Nfp,Nfaces,K)
    vmapM = vmapM.T.ravel()
    mapB = np.nonzero(vmapP == vmapM)[0] + 1
    vmapB = vmapM[mapB - 1]
    mapI, mapO, vmapI, vmapO = 1, K * Nfaces, 1, K * Np       # :43
    return vmapM, vmapP, vmapB, mapB, mapI, mapO, vmapI, vmapO


def StartUp1D(N, VX, EToV):
    """utils/StartUp1D.m:5-39 -- returns all 'globals' as a namespace."""
    g = SimpleNamespace()
    g.N, g.Np, g.Nfp, g.Nfaces = N, N + 1, 1, 2
    g.VX = np.asarray(VX, dtype=float)
    g.K = K = EToV.shape[0]
    g.r = JacobiGL(0, 0, N)                                    # :9
    g.V = Vandermonde1D(N, g.r)                                # :12
    g.invV = np.linalg.inv(g.V)
    g.Dr = Dmatrix1D(N, g.r, g.V)                              # :13
    g.LIFT = Lift1D(g.Np, g.Nfaces, g.Nfp, g.V)                # :16
    va, vb = EToV[:, 0], EToV[:, 1]                            # :19-20
    g.x = np.ones((N + 1, 1)) * g.VX[va][None, :] + 0.5 * (g.r[:, None] + 1) * (g.VX[vb] - g.VX[va])[None, :]
    g.rx, g.J = GeometricFactors1D(g.x, g.Dr)                  # :23
    fmask1 = np.nonzero(np.abs(g.r + 1) < NODETOL)[0]          # :26-28
    fmask2 = np.nonzero(np.abs(g.r - 1) < NODETOL)[0]
    g.Fmask = np.array([fmask1[0], fmask2[0]])
    g.Fx = g.x[g.Fmask, :]                                     # :29
    g.nx = Normals1D(K)                                        # :32
    g.Fscale = 1.0 / g.J[g.Fmask, :]                           # :33
    g.EToV = EToV
    g.EToE, g.EToF = Connect1D(EToV)                           # :36
    (g.vmapM, g.vmapP, g.vmapB, g.mapB, g.mapI, g.mapO, g.vmapI, g.vmapO) = BuildMaps1D(
        g.Np, K, g.Fmask, g.EToE, g.EToF, g.x)                 # :39
    g.rk4a, g.rk4b, g.rk4c = rk4a, rk4b, rk4c
    return g


def startup_uniform(N, xmin, xmax, K):
    """MeshGen1D + StartUp1D, the prologue of utils/One_code.mlx."""
    _, VX, _, EToV = MeshGen1D(xmin, xmax, K)
    return StartUp1D(N, VX, EToV)


def startup_mesh(N, VX):
    """StartUp1D on an arbitrary (non-uniform) vertex list VX[K+1] (refined meshes)."""
    VX = np.asarray(VX, dtype=float)
    K = VX.size - 1
    EToV = np.stack([np.arange(K), np.arange(1, K + 1)], axis=1)
    return StartUp1D(N, VX, EToV)


def fem_setup(n, k, tspan, n_gq):
    """matlab/fem_setup.m:1-41 -- StartUp1D on tspan, then Gauss quadrature + Phi.

    Quirk C-7: the reference overwrites the global `r` (LGL nodes) with the Gauss
    points (:27); we return both (`r` = Gauss points, `r_lgl` = LGL nodes).
    """
    _, VX, _, EToV = MeshGen1D(tspan[0], tspan[1], k)          # :8-23
    g = StartUp1D(n, VX, EToV)                                 # :25
    g.r_lgl = g.r
    g.r, g.w = JacobiGQ(0, 0, n_gq)                            # :27
    n_r = g.r.size
    invVT = np.linalg.inv(g.V.T)                               # :31
    Phi = np.zeros((n_r, g.Np))
    for kq in range(n_r):                                      # :32-39
        for i in range(g.Np):
            p = np.zeros(g.Np)
            for nn in range(g.Np):
                p[nn] = invVT[i, nn] * JacobiP(g.r[kq], 0, 0, nn)[0]
            Phi[kq, i] = p.sum()
    g.Phi = Phi
    return g


def mass_matrix(V):
    """M = (V V')^-1 on the reference element (used as inv(V*V') in dg_march.m:15)."""
    return np.linalg.inv(V @ V.T)


def prolongation(N_from, N_to):
    """Nodal prolongation order N_from -> N_to >= N_from on the reference element:
    P = V_to[:, :Np_from] * inv(V_from)  (build-specified)."""
    r_from = JacobiGL(0, 0, N_from)
    r_to = JacobiGL(0, 0, N_to)
    V_from = Vandermonde1D(N_from, r_from)
    V_to_low = Vandermonde1D(N_from, r_to)                     # modes 0..N_from at the fine nodes
    return V_to_low @ np.linalg.inv(V_from)
