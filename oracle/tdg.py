"""Oracle: DG-in-time ODE solver, reverse-time DG adjoint and per-element adjoint-weighted
indicator -- bug-for-bug NumPy restatement of matlab/dg_march.m, matlab/adj_march.m (both
branches), matlab/adj_rec.m and the semantics of matlab/err_contribution.m (TEST INFRASTRUCTURE, see
oracle/__init__.py).  Batched over the initial value y0 (shared mesh `times`, shared orders),
the batch pattern of python/Main_variable_params.py:330-339.

Parity status: weakly pinned by the reference -- the figure init_nonlin.png (3 digits,
SURVEY App. B.2) and the closed forms of python/factory.py:102-131; the KATs of SURVEY
App. B.2 / B.4 are reproduced in tests/test_oracle_golden.py.  Reference quirks kept
(SURVEY App. C): C-3 (adj_march uses hk = x(1)-x(end) < 0, so its quadrature points lie in the
mirrored interval and the primal polynomial is extrapolated there), C-5 (interpolation by
polyfit/polyval in physical time), C-6 (Newton stops on ||dU|| <= 1e-7), C-7, and adj_march's
hard-coded y0 = 1 (adj_march.m:9).
"""
from __future__ import annotations

import numpy as np

from . import operators as ops


def _polyfit_matrix(x_from, deg, x_to):
    """Matrix of  polyval(polyfit(x_from, U, deg), x_to)  (linear in U): column j is the
    operation applied to the j-th unit vector (quirk C-5: monomial basis in physical time)."""
    n = len(x_from)
    M = np.zeros((len(x_to), n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        M[:, j] = np.polyval(np.polyfit(x_from, e, deg), x_to)
    return M


def primal_element(N, tspan, linear=False):
    """Per-element constants of dg_march.m (shared by the batch)."""
    n_gq = 1 if linear else 30 * N                              # dg_march.m:13 / :29
    g = ops.fem_setup(N, 1, tspan, n_gq)
    x = g.x[:, 0]
    hk = x[-1] - x[0]                                           # :14 / :30
    Minv = ops.mass_matrix(g.V)                                 # inv(V*V')
    S = Minv @ g.Dr                                             # :16 / :53  ((V*V')\Dr = inv(V*V')*Dr... see note)
    Np = g.Np
    out = dict(g=g, x=x, hk=hk, Np=Np)
    if linear:
        B = np.zeros((Np, Np)); B[-1, -1] = 1.0                 # :17
        out["A"] = -S.T + B - hk / 2 * Minv                     # :15,:18
    else:
        B = np.zeros((Np, Np)); B[-1, -1] = -1.0                # :54
        out["A"] = S.T + B                                      # :57
        x_interp = x[0] + (1 + g.r) * hk / 2                    # :48
        out["Iq"] = _polyfit_matrix(x, N, x_interp)             # :47-49 as a matrix
        out["Phi"], out["w"] = g.Phi, g.w
    return out


def dg_march(Ns, Ks, times, y0, linear=False, tol=1e-7, maxit=500):
    """matlab/dg_march.m:1-80.  y0: scalar or (B,).  Returns (t, y, its): lists over elements of
    node times (Np,), nodal values (B, Np) and Newton iteration counts (B,)."""
    y0 = np.atleast_1d(np.asarray(y0, dtype=float))
    B = y0.size
    uR = y0.copy()
    t, y, its = [], [], []
    for k in range(Ks):
        el = primal_element(int(Ns[k]), (times[k], times[k + 1]), linear)
        Np, A, hk = el["Np"], el["A"], el["hk"]
        if linear:
            F = np.zeros((B, Np)); F[:, 0] = uR                 # :19
            U = np.linalg.solve(A, F.T).T                       # :21
            its.append(np.ones(B, dtype=int))
        else:
            Phi, w, Iq = el["Phi"], el["w"], el["Iq"]
            U = np.repeat(uR[:, None], Np, axis=1)              # :40
            it = np.zeros(B, dtype=int)
            active = np.ones(B, dtype=bool)
            err = np.ones(B)
            while True:
                active = (it <= maxit) & (err > tol)            # :44
                if not active.any():
                    break
                Ua = U[active]
                ur = Ua @ Iq.T                                  # :47-49
                Mt = hk / 2 * (w * np.sin(ur)) @ Phi            # :51,:53   Phi'*(w.*sin)
                dM = hk / 2 * np.einsum("qi,bq,qj->bij", Phi, w * np.cos(ur), Phi)   # :52,:54
                F = np.zeros_like(Ua); F[:, 0] = uR[active]     # :56
                R = Ua @ A.T + Mt + F                           # :61
                dU = np.linalg.solve(A[None] + dM, R[..., None])[..., 0]             # :62
                Un = Ua - dU                                    # :63
                err[active] = np.linalg.norm(Ua - Un, axis=1)   # :64
                U[active] = Un
                it[active] += 1
            its.append(it)
        uR = U[:, -1].copy()                                    # :23 / :74
        t.append(el["x"]); y.append(U)
    return t, y, its


def adjoint_element(Na, tk_primal, linear=False, quirk_c3=True):
    """Per-element constants of adj_march.m for adjoint order Na on the element whose primal
    nodes are tk_primal.  quirk_c3=False evaluates the primal at the quadrature points INSIDE the
    element instead of the mirrored interval the reference uses (adj_march.m:78 with hk < 0);
    everything else is unchanged."""
    tspan = (tk_primal[0], tk_primal[-1])
    g = ops.fem_setup(Na, 1, tspan, 1 if linear else 2 * Na)    # adj_march.m:17 / :71
    x = g.x[:, 0]
    hk = x[0] - x[-1]                                           # :18 / :72  (negative: quirk C-3)
    Minv = ops.mass_matrix(g.V)
    S = Minv @ g.Dr
    Np = g.Np
    deg = len(tk_primal) - 1                                    # :36 length(y1{k})-1 ; :75 Ns(k)-1
    out = dict(g=g, x=x, hk=hk, Np=Np, S=S, Minv=Minv)
    out["Ix"] = _polyfit_matrix(tk_primal, deg, x)              # uh_k = polyval(pu, x)
    if not linear:
        r_interp = tk_primal[0] + (1 + g.r) * (hk if quirk_c3 else -hk) / 2    # :78 (mirrored interval)
        out["Iq"] = _polyfit_matrix(tk_primal, deg, r_interp)
        out["Phi"], out["w"] = g.Phi, g.w
    return out


def adj_march(Ns, Ks, times, y1, t1, linear=False, y0_hard=1.0, quirk_c3=True):
    """matlab/adj_march.m:1-122.  Ns = adjoint orders (MAIN.m:34 passes Ns+1); y1 / t1 = primal
    from dg_march (lists over elements; y1[k] is (B, Np_primal)).  Returns (t, v, err) with
    v[k] (B, Np) and err (B, Ks) -- signed; MAIN.m:51 takes abs."""
    B = y1[0].shape[0]
    vL = np.zeros(B)
    t, v = [None] * Ks, [None] * Ks
    err = np.zeros((B, Ks))
    for k in range(Ks - 1, -1, -1):
        el = adjoint_element(int(Ns[k]), t1[k], linear, quirk_c3)
        Np, S, Minv, hk = el["Np"], el["S"], el["Minv"], el["hk"]
        uh = y1[k] @ el["Ix"].T                                 # primal at the adjoint nodes
        F0 = np.zeros((B, Np))
        F0[:, 0] = y0_hard if k == 0 else y1[k - 1][:, -1]      # :42-46 / :108-112 (y0 = 1, :9)
        if linear:
            M = hk / 2 * Minv                                   # :19
            m = np.zeros((Np, Np)); m[0, 0] = -1.0              # :21
            A = -S.T + m - M                                    # :22
            F = np.tile(M @ np.ones(Np), (B, 1)); F[:, -1] -= vL    # :28
            vk = np.linalg.solve(A, F.T).T                      # :31
            m2 = m.copy(); m2[0, 0] = 0.0; m2[-1, -1] = 1.0     # :40  m([1,end]) = [0,1]
            A2 = -S.T + m2 + M                                  # :41
            err[:, k] = np.einsum("bi,bi->b", vk, -(uh @ A2.T) + F0)     # :47
        else:
            Phi, w = el["Phi"], el["w"]
            ur = y1[k] @ el["Iq"].T                             # :79
            Mv = hk / 2 * np.einsum("qi,bq,qj->bij", Phi, w * np.cos(ur), Phi)   # :81-82
            Mk = hk / 2 * Minv                                  # :83
            Bm = np.zeros((Np, Np)); Bm[0, 0] = -1.0            # :85
            A = (-S.T + Bm)[None] - Mv                          # :86
            F = np.tile(Mk @ np.ones(Np), (B, 1)); F[:, -1] -= vL   # :96
            vk = np.linalg.solve(A, F[..., None])[..., 0]       # :98
            Mt = hk / 2 * (w * np.sin(ur)) @ Phi                # :104-105
            B2 = np.zeros((Np, Np)); B2[-1, -1] = -1.0          # :107 (the store at :103 is dead, C-4)
            A2 = -S.T - B2                                      # :115
            err[:, k] = np.einsum("bi,bi->b", vk, -(uh @ A2.T) - Mt + F0)   # :117
        vL = vk[:, 0].copy()                                    # :33 / :100
        t[k] = el["x"]; v[k] = vk
    return t, v, err


# utils/Globals1D.m:37-42 -- left Radau points as the reference tabulates them: exact for
# m <= 3, SIX-DIGIT decimals for m = 4, 5 (kept as written: bug-for-bug)
RADAU = {
    1: np.array([-1.0]),
    2: np.array([-1.0, 1.0 / 3.0]),
    3: np.array([-1.0, (1.0 - np.sqrt(6.0)) / 5.0, (1.0 + np.sqrt(6.0)) / 5.0]),
    4: np.array([-1.0, -0.575319, 0.181066, 0.822824]),
    5: np.array([-1.0, -0.72048, -0.167181, 0.446314, 0.885792]),
}


def adjrec_element(N, tk_primal):
    """Per-element constants of the linear branch of adj_rec.m (matlab/adj_rec.m:18-71)."""
    tspan = (tk_primal[0], tk_primal[-1])
    g = ops.fem_setup(N, 1, tspan, 1)                            # :20
    x = g.x[:, 0]
    hk = x[0] - x[-1]                                           # :21 (negative, as in adj_march)
    Minv = ops.mass_matrix(g.V)
    M = hk / 2 * Minv                                           # :22
    S = Minv @ g.Dr                                             # :23
    Np = g.Np
    m = np.zeros((Np, Np)); m[0, 0] = -1.0                      # :24
    rad_m = N + 1                                               # :36
    rad_x = tspan[0] + (1 + RADAU[rad_m]) * abs(hk) / 2         # :37-38
    x_rec = np.concatenate([rad_x, [tspan[1]]])                 # :46
    ge = ops.fem_setup(rad_m, 1, tspan, 1)                      # :50  (hk is NOT recomputed)
    xe = ge.x[:, 0]
    Minv_e = ops.mass_matrix(ge.V)
    Me = hk / 2 * Minv_e                                        # :51
    Se = Minv_e @ ge.Dr                                         # :52
    me = np.zeros((ge.Np, ge.Np)); me[-1, -1] = 1.0             # :53
    return dict(Np=Np, x_rec=x_rec,
                A=-S.T + m - M,                                 # :25
                f1=M @ np.ones(Np),                             # :31
                R=_polyfit_matrix(x, Np - 1, rad_x),            # :42-44  v_s -> Radau points
                H=_polyfit_matrix(x_rec, rad_m, xe),            # :47-48,:65  v_rec -> enriched nodes
                A2=-Se.T + me + Me,                             # :54
                Ix=_polyfit_matrix(tk_primal, len(tk_primal) - 1, xe))   # :62-64


def adj_rec(Ns, Ks, times, y1, t1, linear=False, y0_hard=1.0):
    """matlab/adj_rec.m:1-88 -- adjoint at the PRIMAL order, reconstructed to order N+1 through
    the Radau points plus the inflow value (disabled in the reference: MAIN.m:35).  Ns = primal
    orders (MAIN.m:35 passes Ns).  linear=True: adj_rec.m:18-71.  linear=False is the branch the
    file ships with (`linear = false`, :11): unfinished -- it assembles a mass matrix per element
    and returns empty cells and err = 0 (:73-87); restated as exactly that.
    Returns (t, v, err): t[k] = [Radau points; t_{k+1}], v[k] (B, N_k+2), err (B, Ks) signed."""
    B = y1[0].shape[0]
    t, v = [None] * Ks, [None] * Ks
    err = np.zeros((B, Ks))
    if not linear:
        return t, v, err
    vL = np.zeros(B)
    for k in range(Ks - 1, -1, -1):
        el = adjrec_element(int(Ns[k]), t1[k])
        F = np.tile(el["f1"], (B, 1)); F[:, -1] -= vL           # :31
        vs = np.linalg.solve(el["A"], F.T).T                    # :40
        vrec = np.concatenate([vs @ el["R"].T, vL[:, None]], axis=1)     # :44
        vh = vrec @ el["H"].T                                   # :65
        uh = y1[k] @ el["Ix"].T                                 # :64
        F0 = np.zeros_like(uh)
        F0[:, 0] = y0_hard if k == 0 else y1[k - 1][:, -1]      # :55-59 (y0 = 1, :9)
        err[:, k] = np.einsum("bi,bi->b", vh, -(uh @ el["A2"].T) + F0)   # :66
        v[k] = vrec                                             # :68
        vL = vrec[:, 0].copy()                                  # :69
        t[k] = el["x_rec"]                                      # :70
    return t, v, err


def err_contribution_linear_exact(Ks, Ns, uh, t1, nquad=64):
    """Semantics of matlab/err_contribution.m:21-43 (unused by MAIN.m:50): err_i = int over
    element i of a(t) (u_h(t) - u_h'(t)) dt with the exact adjoint of a' = -a - 1, a(1) = 0,
    i.e. a(t) = e^{1-t} - 1, plus u(1) - 1 on the first element (:42-43).  MATLAB's adaptive
    `integral` is replaced by Gauss quadrature (the integrand is a polynomial times exp)."""
    xq, wq = np.polynomial.legendre.leggauss(nquad)
    B = uh[0].shape[0]
    err = np.zeros((B, Ks))
    for i in range(Ks):
        tu = t1[i]
        a, b = tu[0], tu[-1]
        tq = 0.5 * (b - a) * xq + 0.5 * (a + b)
        adj = np.exp(1.0 - tq) - 1.0
        for bb in range(B):
            pu = np.polyfit(tu, uh[i][bb], int(Ns[i]))
            res = np.polyval(pu, tq) - np.polyval(np.polyder(pu), tq)
            err[bb, i] = 0.5 * (b - a) * np.sum(wq * adj * res)
        if i == 0:
            err[:, i] += uh[i][:, 0] - 1.0
    return err


def refine(times, Ns, err, n):
    """matlab/MAIN.m:137-141: refine the element with the largest |err| by midpoint insertion
    (lowest index on ties -- SURVEY quirk C-10)."""
    ref_i = int(np.argmax(np.abs(err)))
    times = np.insert(np.asarray(times, float), ref_i + 1, 0.5 * (times[ref_i] + times[ref_i + 1]))
    Ns = np.append(np.asarray(Ns), n)
    return times, Ns, ref_i
