"""CPU tests of the product's host side: the galerkin.py mirror against the oracle, the C
library's host code (even/odd operator blocks) through an emulation of the kernel's algebra,
the C-ABI surface, and the multi-GPU plumbing on gloo with world_size 2."""
import ctypes as C
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import eo_emulator as eo
import modal_emulator as em
from oracle import advec
from oracle import operators as ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- galerkin.py mirror
@pytest.mark.parametrize("N", range(1, 10))
def test_basegalerkin_matches_oracle(pkg, N):
    g = pkg.BaseGalerkin1D(n=N, k=7, domain=(0.5, 3.0), n_gq=2 * N)
    o = ops.fem_setup(N, 7, (0.5, 3.0), 2 * N)
    pairs = [(g.r_lgl, o.r_lgl), (g.v, o.V), (g.inv_v, o.invV), (g.d_r, o.Dr), (g.lift, o.LIFT), (g.x, o.x),
             (g.r_x, o.rx), (g.j_mat, o.J), (g.f_x, o.Fx), (g.n_x, o.nx), (g.f_scale, o.Fscale), (g.v_x, o.VX),
             (g.r, o.r), (g.w, o.w), (g.phi, o.Phi), (g.mass, ops.mass_matrix(o.V))]
    for a, b in pairs:
        assert a.shape == b.shape
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-10)
    assert np.array_equal(g.e_to_e, o.EToE) and np.array_equal(g.e_to_f, o.EToF)
    assert np.array_equal(g.e_to_v, o.EToV)
    np.testing.assert_allclose(g.prolongation_to(pkg.BaseGalerkin1D(n=N + 1, k=7, domain=(0.5, 3.0))),
                               ops.prolongation(N, N + 1), atol=1e-11)


def test_basegalerkin_interface(pkg):
    """Attribute names / shapes / defaults of python/galerkin.py:18-24,199-263."""
    g = pkg.BaseGalerkin1D()
    assert (g.n, g.k, g.n_gq, g.n_fp, g.n_faces, g.node_tol) == (1, 2, 2, 1, 2, 1e-10)
    assert list(g.domain) == [0.0, 1.0]
    for name in ("n_p r v inv_v d_r lift x r_x j_mat f_mask f_x n_x f_scale e_to_e e_to_f v_map_m v_map_p "
                 "v_map_b map_b map_i map_o v_map_i v_map_o v_x e_to_v w n_r phi").split():
        assert hasattr(g, name), name
    for name in ("jacobiGQ jacobiGL jacobiP vandermonde1D gradJacobiP gradVandermonde1D dMatrix1D lift1D "
                 "geometricFactors1D normals1D connect1D buildMaps1D startUp1D").split():
        assert callable(getattr(g, name)), name
    g = pkg.BaseGalerkin1D(n=3, k=5)
    assert g.x.shape == (4, 5) and g.r_x.shape == (4, 5) and g.f_scale.shape == (2, 5) and g.phi.shape == (3, 4)
    # maps: 0-based into the row-major (Np, K) flattening; interior faces meet the neighbour's node
    xf = g.x.ravel()
    assert np.allclose(xf[g.v_map_m[1, :-1]], xf[g.v_map_p[1, :-1]])
    assert g.v_map_p[1, 0] == g.v_map_m[0, 1]
    assert (g.map_i, g.map_o, g.v_map_i, g.v_map_o) == (0, 9, 0, 19)
    assert g.map_b.tolist() == [0, 9]


def test_nonuniform_mesh(pkg):
    vx = np.array([0.0, 0.25, 0.5, 1.0, 2.0])
    g = pkg.BaseGalerkin1D(n=2, v_x=vx)
    o = ops.startup_mesh(2, vx)
    np.testing.assert_allclose(g.r_x, o.rx, rtol=1e-13)
    np.testing.assert_allclose(g.f_scale, o.Fscale, rtol=1e-13)
    assert g.k == 4


# ---------------------------------------------------------------- C host code + kernel algebra
CASES = [(4, 10, "periodic", 0.0), (4, 10, "inflow", 1.0), (3, 7, "inflow", 0.3), (8, 16, "periodic", 0.0),
         (1, 5, "periodic", 1.0), (2, 20, "inflow", 1.0), (7, 6, "inflow", 0.0), (5, 8, "periodic", 0.5)]


@pytest.mark.parametrize("N,K,bc,alpha", CASES)
def test_modal_algebra_matches_oracle(pkg, lib, N, K, bc, alpha):
    """The kernel's formulation (modal basis, parity-sparse derivative, stage scalings, injection
    as prolongation) on the operators the C library's host code builds, against the oracle."""
    gc = pkg.BaseGalerkin1D(n=N, k=K, domain=(0, 2 * math.pi))
    gf = pkg.BaseGalerkin1D(n=N + 1, k=K, domain=(0, 2 * math.pi))
    oc = ops.startup_uniform(N, 0, 2 * math.pi, K)
    of = ops.startup_uniform(N + 1, 0, 2 * math.pi, K)
    a = 2 * math.pi
    dt, S = advec.cfl_dt(oc, 0.15)
    rng = np.random.default_rng(N * 100 + K)
    u0 = np.sin(oc.x) + 0.1 * rng.standard_normal(oc.x.shape)
    per = bc == "periodic"
    for g in (oc, of):          # the kernel's per-element rx (mean over the element's nodes)
        g.rx = np.broadcast_to(g.rx.sum(axis=0, keepdims=True) / g.rx.shape[0], g.rx.shape).copy()
    ref = advec.fwd_adj_indicator(u0, oc, of, a, dt, S, alpha=alpha, bc=bc, inflow=advec.INFLOW_SIN_AT)
    rk = (ops.rk4a, ops.rk4b, ops.rk4c)
    out = em.fused(lib, gc, gf, u0, a, dt, S, alpha, per, rk, gc.quad_weights(), gf.quad_weights(),
                   inflow_fn=None if per else (lambda t: -np.sin(a * t)))
    assert out["viol"] < 1e-12
    np.testing.assert_allclose(out["uT"], ref["uT"], rtol=0, atol=1e-12 * np.max(np.abs(ref["uT"])))
    np.testing.assert_allclose(out["lam0"], ref["lam0"], rtol=0, atol=1e-12 * np.max(np.abs(ref["lam0"])))
    assert abs(out["J"] - ref["J"]) <= 1e-12 * max(1.0, abs(ref["J"]))
    assert np.all(np.abs(out["eta"] - ref["eta"]) <= 1e-12 * ref["eta_scale"])


@pytest.mark.parametrize("bc", ["periodic", "inflow"])
def test_hp_mode_masks_are_the_hp_scheme(pkg, lib, bc):
    """The claim behind `dgadj_set_element_orders` checked without a GPU: the padded modal march with the HP kernels'
    mode masks (tests/modal_emulator.py, on the operators of the C library's host code) IS the hp scheme -- every
    element with the StartUp1D operators of its own order, ragged arrays, a dense global matrix (oracle/advec_hp.py):
    terminal state, functional, the adjoint as a covector of each element's own enriched space, the indicators."""
    from oracle import advec_hp as ohp
    N, orders = 4, [2, 4, 3, 4, 1, 2, 4, 1]
    vx = np.array([0.0, 0.5, 1.1, 1.6, 2.4, 3.0, 3.9, 4.6, 2 * math.pi])
    K = len(orders)
    gc, gf = pkg.BaseGalerkin1D(n=N, v_x=vx), pkg.BaseGalerkin1D(n=N + 1, v_x=vx)
    c = ohp.HpSpace(orders, vx)
    rng = np.random.default_rng(17)
    u0r = np.concatenate([np.sin(x + 0.7) + 0.4 * rng.standard_normal(x.size) for x in c.x])
    u0p = ohp.pad(c, u0r, N)
    a, dt, S = 1.3, 2e-3, 40
    per = bc == "periodic"
    alpha = 0.0 if per else 0.3
    ref = ohp.fwd_adj_indicator(u0r, orders, vx, a, dt, S, alpha, per)
    rk = (ops.rk4a, ops.rk4b, ops.rk4c)
    out = em.fused(lib, gc, gf, u0p, a, dt, S, alpha, per, rk, gc.quad_weights(), gf.quad_weights(),
                   nodes_per_element=np.asarray(orders) + 1)
    f = ref["spaces"][1]
    scale = np.max(np.abs(ref["uT"]))
    assert np.max(np.abs(ohp.unpad(c, out["uT"], N) - ref["uT"])) < 1e-12 * scale
    assert np.max(np.abs(ohp.pad(c, ohp.unpad(c, out["uT"], N), N) - out["uT"])) < 1e-13 * scale   # stays in the spaces
    assert abs(out["J"] - ref["J"]) < 1e-12 * max(1.0, abs(ref["J"]))
    assert np.max(np.abs(ohp.unpad_covector(f, out["lam0"], N + 1) - ref["lam0"])) < 1e-12 * np.max(np.abs(ref["lam0"]))
    assert np.max(np.abs(out["eta"] - ref["eta"]) / ref["eta_scale"]) < 1e-12


def test_modal_operator_structure(pkg, lib):
    """V^-1 Dr V is strictly upper triangular and parity sparse (floor(Np^2/4) non-zeros), the
    lift is V^T E, the prolongation is the injection; a wrong V is reported."""
    for N in range(1, 10):
        g = pkg.BaseGalerkin1D(n=N, k=3)
        o = em.modal_ops(lib, g)
        Np = N + 1
        assert o["viol"] < 1e-12
        assert np.count_nonzero(o["D"]) == (Np * Np) // 4
        np.testing.assert_allclose(o["V"] @ o["D"] @ o["iV"], g.d_r, atol=1e-11)
        np.testing.assert_allclose(o["p"], g.v[-1, :], rtol=0, atol=0)
        gf = pkg.BaseGalerkin1D(n=N + 1, k=3)
        Ph = np.linalg.inv(gf.v) @ g.prolongation_to(gf) @ g.v
        np.testing.assert_allclose(Ph, np.eye(Np + 1, Np), atol=1e-12)
    g = pkg.BaseGalerkin1D(n=4, k=3)
    Dnz, p, iV = np.zeros(26), np.zeros(5), np.zeros((5, 5))
    v = C.c_double()
    ptr = lambda a: C.c_void_p(a.ctypes.data)
    Vbad = np.ascontiguousarray(g.v[:, ::-1])             # permuted modes: not the Legendre ordering
    Dr, LIFT = np.ascontiguousarray(g.d_r), np.ascontiguousarray(g.lift)
    assert lib.dgadj_host_modal_operators(5, ptr(Dr), ptr(LIFT), ptr(Vbad), ptr(Dnz), ptr(p), ptr(iV), C.byref(v)) == 0
    assert v.value > 1e-3                                 # dgadj_set_operators refuses it (> 1e-9)


def test_even_odd_blocks_reproduce_the_nodal_derivative(pkg, lib):
    """dgadj_host_eo_operators (used by the Burgers kernels): the half-size blocks applied to the
    symmetric / antisymmetric parts give Dr u and LIFT g."""
    for N in (1, 4, 7, 8):
        g = pkg.BaseGalerkin1D(n=N, k=2)
        b = eo.eo_blocks(lib, g)
        assert b["viol"] < 1e-13
        rng = np.random.default_rng(N)
        u = rng.standard_normal((N + 1, 3))
        e, o = eo.to_eo(u)
        g2 = rng.standard_normal((2, 3))
        E = b["DE"] @ o + b["LS"][:, None] * (g2[0] + g2[1])
        O = b["DO"] @ e + b["LA"][:, None] * (g2[0] - g2[1])
        np.testing.assert_allclose(eo.from_eo(E, O, 0.5), g.d_r @ u + g.lift @ g2, atol=1e-11)


def test_transposed_volume_term_through_the_even_odd_blocks(pkg, lib):
    """The fused Burgers kernel transposes its volume term through the SAME even / odd blocks the forward term uses
    (csrc/dgadj_burgers_fused.cu, BG_TRANSPOSED_EO): the forward form as the kernel computes it equals 2 Dr F, and the
    transposed form is its exact transpose (dot-product identity at rounding) -- for even and odd node counts."""
    for N in range(1, 9):
        g = pkg.BaseGalerkin1D(n=N, k=2)
        b = eo.eo_blocks(lib, g)
        DE, DO, HE, HO, Np = b["DE"], b["DO"], b["HE"], b["HO"], N + 1
        rng = np.random.default_rng(N)

        def forward(F):                          # bg_stage: fe / fo, E = DE fo, O = DO fe, rhs_i = E + O, rhs_{N-i} = E - O
            fe = np.array([F[i] + F[Np - 1 - i] for i in range(HO)] + ([F[HO]] if Np & 1 else []))
            fo = np.array([F[i] - F[Np - 1 - i] for i in range(HO)])
            E, O = DE @ fo, DO @ fe
            rhs = np.zeros(Np)
            for i in range(HO):
                rhs[i], rhs[Np - 1 - i] = E[i] + O[i], E[i] - O[i]
            if Np & 1:
                rhs[HO] = E[HO] + E[HO]
            return rhs

        def transposed(w):                       # the reverse stage: we / wo, A = DE^T we, B = DO^T wo
            we = np.array([w[i] + w[Np - 1 - i] for i in range(HO)] + ([2.0 * w[HO]] if Np & 1 else []))
            wo = np.array([w[i] - w[Np - 1 - i] for i in range(HO)])
            A, B = DE.T @ we, DO.T @ wo
            acc = np.zeros(Np)
            for j in range(HO):
                acc[j], acc[Np - 1 - j] = A[j] + B[j], B[j] - A[j]
            if Np & 1:
                acc[HO] = B[HO]
            return acc

        for _ in range(5):
            F, w = rng.standard_normal(Np), rng.standard_normal(Np)
            np.testing.assert_allclose(forward(F), 2.0 * g.d_r @ F, atol=1e-11)
            assert abs(transposed(w) @ F - w @ forward(F)) < 1e-12 * (np.abs(w) @ np.abs(forward(F)) + 1.0)
            np.testing.assert_allclose(transposed(w), 2.0 * g.d_r.T @ w, atol=1e-11)


def test_eo_rejects_asymmetric_operator(lib):
    Np = 4
    Dr = np.arange(16, dtype=float).reshape(4, 4)
    LIFT = np.ones((4, 2))
    DE, DO, LS, LA = np.zeros(25), np.zeros(25), np.zeros(5), np.zeros(5)
    v = C.c_double()
    p = lambda a: C.c_void_p(a.ctypes.data)
    assert lib.dgadj_host_eo_operators(Np, p(Dr), p(LIFT), p(DE), p(DO), p(LS), p(LA), C.byref(v)) == 0
    assert v.value > 1e-3        # dgadj_set_operators refuses such a set (violation > 1e-10)
    assert lib.dgadj_host_eo_operators(1, p(Dr), p(LIFT), p(DE), p(DO), p(LS), p(LA), C.byref(v)) == -1


# ---------------------------------------------------------------- C-ABI surface
def test_abi_exports_every_declared_symbol(pkg, lib):
    with open(os.path.join(ROOT, "include", "dgadj.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b(dgadj_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 24
    assert declared == set(pkg._lib.PROTOTYPES), declared ^ set(pkg._lib.PROTOTYPES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.dgadj_version() == 100
    assert C.sizeof(pkg._lib.Config) == 40 and C.sizeof(pkg._lib.MarchArgs) == 56


def test_create_argument_checks_and_no_cpu_fallback(pkg, lib):
    h = C.c_void_p(0)
    cfg = pkg._lib.Config(device=0, N=4, K=10, bc=0, inflow=1, functional=0, scheme=0, reserved=0, alpha=1.0)
    assert lib.dgadj_create(None, C.byref(h)) == pkg._lib.ERR_INVALID
    bad = pkg._lib.Config(device=0, N=10, K=10, bc=0, inflow=1, functional=0, scheme=0, reserved=0, alpha=1.0)
    assert lib.dgadj_create(C.byref(bad), C.byref(h)) == pkg._lib.ERR_UNSUPPORTED
    bad = pkg._lib.Config(device=0, N=4, K=4096, bc=0, inflow=1, functional=0, scheme=0, reserved=0, alpha=1.0)
    assert lib.dgadj_create(C.byref(bad), C.byref(h)) == pkg._lib.ERR_UNSUPPORTED
    bad = pkg._lib.Config(device=0, N=4, K=10, bc=7, inflow=1, functional=0, scheme=0, reserved=0, alpha=1.0)
    assert lib.dgadj_create(C.byref(bad), C.byref(h)) == pkg._lib.ERR_INVALID
    import torch
    if not torch.cuda.is_available():
        assert lib.dgadj_create(C.byref(cfg), C.byref(h)) == pkg._lib.ERR_NO_DEVICE
        with pytest.raises(pkg.DgadjError):
            pkg.AdvecDG1D(4, 10)


# ---------------------------------------------------------------- sharding + gloo world_size 2
def test_shard_range(pkg):
    for B, W in [(10, 3), (65536, 8), (5, 8), (1, 1)]:
        parts = [pkg.shard_range(B, r, W) for r in range(W)]
        assert parts[0][0] == 0 and parts[-1][1] == B
        assert all(parts[i][1] == parts[i + 1][0] for i in range(W - 1))
        assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


def test_batch_mean_refine(pkg):
    sums = np.array([3.0, 9.0, 9.0, 1.0, 22.0, 0.0, 9.0, 0.0])
    mean, idx = pkg.batch_mean_refine(sums, 3)
    assert idx == 1 and mean.tolist() == [1.0, 3.0, 3.0, 1.0 / 3.0]


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import dgadj_loader
pkg = dgadj_loader.load_package()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
K, B = 6, 11
rng = np.random.default_rng(0)
eta = rng.standard_normal((B, K)); J = rng.standard_normal(B)
lo, hi = pkg.shard_range(B, rank, world)
e, j = eta[lo:hi], J[lo:hi]
part = torch.tensor(np.concatenate([np.abs(e).sum(0), [np.abs(e).sum(), (e**2).sum(), np.abs(e).max(), j.sum()]]))
tot = pkg.allreduce_indicators(part, ordered=True)
tot2 = pkg.allreduce_indicators(part, ordered=False)
full = np.concatenate([np.abs(eta).sum(0), [np.abs(eta).sum(), (eta**2).sum(), np.abs(eta).max(), J.sum()]])
assert np.allclose(tot.numpy(), full, rtol=1e-13), (tot, full)
assert np.allclose(tot2.numpy(), full, rtol=1e-13)
gathered = [torch.empty_like(tot) for _ in range(world)]
dist.all_gather(gathered, tot)
assert all(torch.equal(gathered[0], g) for g in gathered)      # bit-identical on every rank
mean, idx = pkg.batch_mean_refine(tot, B)
assert idx == int(np.argmax(np.abs(eta).sum(0)))
full_eta = pkg.gather_indicators(torch.tensor(e))           # ragged slices (6 and 5 rows)
assert torch.equal(full_eta, torch.tensor(eta))
# count-independent form: block partials (4 trajectories per block here) combined in global block order
# give the same bits as the one-rank combination of all blocks, whatever the number of ranks
def block_parts(e, j, R):
    rows = []
    for b0 in range(0, len(e), R):
        ee, jj = e[b0:b0 + R], j[b0:b0 + R]
        rows.append(np.concatenate([np.abs(ee).sum(0), [np.abs(ee).sum(), (ee**2).sum(), np.abs(ee).max(), jj.sum()]]))
    return torch.tensor(np.stack(rows))
B2 = 16
eta2 = rng.standard_normal((B2, K)); J2 = rng.standard_normal(B2)
lo, hi = pkg.shard_range(B2, rank, world)          # 8 + 8: whole blocks per rank
mine = pkg.allreduce_indicator_blocks(block_parts(eta2[lo:hi], J2[lo:hi], 4))
one = pkg.combine_blocks(block_parts(eta2, J2, 4))  # what a single rank computes
assert torch.equal(mine, one), (mine, one)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gloo_two_rank_indicator_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


# ---------------------------------------------------------------- MATLAB fixture I/O
def test_save_load_globals_txt_matlab_conventions(pkg, tmp_path):
    """utils/Save_to_1D_global_data.m file set: names, MATLAB index conventions (1-based,
    column-major) -- compared with the oracle's MATLAB-numbered BuildMaps1D and the numbers
    MATLAB embedded in utils/One_code.mlx."""
    import json
    from adjoint_ode_adaptivity_b200 import fixtures
    g = pkg.BaseGalerkin1D(n=2, k=20, domain=(0.0, 1.0))
    o = ops.startup_uniform(2, 0.0, 1.0, 20)
    fixtures.save_globals_txt(g, tmp_path, dt=1.5e-3)
    names = {f[:-4] for f in os.listdir(tmp_path)}
    assert {"Dr", "EToE", "EToF", "Fmask", "Fscale", "Fx", "invV", "J", "K", "LIFT", "mapB", "mapI", "mapO", "N",
            "Nfaces", "Nfp", "NODETOL", "Np", "nx", "r", "rk4a", "rk4b", "rk4c", "rx", "V", "vmapB", "vmapI",
            "vmapM", "vmapO", "vmapP", "VX", "x", "dt"} <= names
    d = fixtures.load_globals_txt(tmp_path)
    assert d["vmapM"].ravel().astype(int).tolist() == o.vmapM.tolist()
    assert d["vmapP"].ravel().astype(int).tolist() == o.vmapP.tolist()
    assert (int(d["mapI"]), int(d["mapO"]), int(d["vmapI"]), int(d["vmapO"])) == (o.mapI, o.mapO, o.vmapI, o.vmapO)
    assert d["vmapB"].ravel().astype(int).tolist() == o.vmapB.tolist() and d["mapB"].ravel().astype(int).tolist() == o.mapB.tolist()
    assert np.array_equal(d["EToE"].astype(int), o.EToE + 1) and np.array_equal(d["EToF"].astype(int), o.EToF + 1)
    np.testing.assert_array_equal(d["Dr"], g.d_r)           # %.17g is lossless
    np.testing.assert_array_equal(d["x"], g.x)
    with open(os.path.join(ROOT, "tests", "golden", "mlx_one_code.json")) as f:
        gold = {(it["name"], it["line"]): it["value"] for it in json.load(f)["items"]}
    np.testing.assert_allclose(d["LIFT"], gold[("LIFT", 148)], atol=5e-5)
    np.testing.assert_allclose(d["rx"], gold[("rx", 149)], atol=5e-3)
    opsd = fixtures.operators_from_globals(d)
    np.testing.assert_allclose(opsd["Mref"], g.mass, rtol=1e-12)


# ---------------------------------------------------------------- a plain-C client of the ABI
def build_c_client(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    pkgdir = os.path.join(ROOT, "adjoint-ode-adaptivity_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", pkgdir, "-ldgadj", "-lm",
                        "-Wl,-rpath," + pkgdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_client_compiles_and_refuses_cpu(tmp_path):
    """include/dgadj.h is valid C99 and the library links from C; without a GPU the client gets
    DGADJ_ERR_NO_DEVICE (exit code 3) -- there is no CPU fallback."""
    import torch
    exe = build_c_client(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if not torch.cuda.is_available():
        assert r.returncode == 3, r.stdout + r.stderr
        assert "create rc -2" in r.stdout


def test_division_free_quotient_has_the_bits_of_the_division(tmp_path):
    """tests/div_rn_check.c: the quotient the kernels form from the correctly rounded reciprocal (the limited cells'
    path of the fused Burgers kernel, the substitutions of the time-DG solves) equals the IEEE division bit for bit;
    the Newton march's stopping rule decided without the square root equals `sqrt(e2) > tol` everywhere."""
    exe = str(tmp_path / "div_rn_check")
    r = subprocess.run(["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-Wall", "-Werror",
                        os.path.join(ROOT, "tests", "div_rn_check.c"), "-o", exe, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "2000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "pairs, 0 differences" in r.stdout and "stopping rule: 0 differences" in r.stdout, r.stdout


def test_division_free_minmod_equals_the_reference_formulation(tmp_path):
    """tests/minmod_check.c: the kernels' minmod (compares and selects only) against utils/minmod.m:6-12
    (sign sum, min of the absolute values), bit for bit on triples rich in zeros, ties and mixed signs; and the warp
    maximum of |u| as two 32-bit reductions on the bit pattern (idle lanes at -1) against the plain maximum."""
    exe = str(tmp_path / "minmod_check")
    r = subprocess.run(["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-Wall", "-Werror",
                        os.path.join(ROOT, "tests", "minmod_check.c"), "-o", exe, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "result): 0 differences" in r.stdout and "bit pattern: 0 differences" in r.stdout, r.stdout


def test_split_dense_solve_equals_the_one_piece_solve(tmp_path):
    """tests/lu_split_check.c: the time-DG kernels' dense solve restated on the host -- the reused pivot reciprocals
    give the bits of the IEEE divisions, and lu_factor + lu_apply (the lane-per-element adjoint kernel) the bits of the
    one-piece solve, on random systems of size 2..7 of which a good share needs row swaps."""
    exe = str(tmp_path / "lu_split_check")
    r = subprocess.run(["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-Wall", "-Werror",
                        os.path.join(ROOT, "tests", "lu_split_check.c"), "-o", exe, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and " 0 differences to the IEEE divisions, 0 between" in r.stdout, r.stdout


def test_missing_library_fails_loudly(tmp_path):
    """No CPU fallback: with libdgadj.so absent the package import raises ImportError."""
    code = ("import sys; sys.path.insert(0, %r); import dgadj_loader\n"
            "try:\n    dgadj_loader.load_package()._lib.load()\nexcept ImportError as e:\n    print('IMPORTERROR', e); sys.exit(7)\n" % ROOT)
    env = dict(os.environ, DGADJ_LIB=str(tmp_path / "nope" / "libdgadj.so"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 7 and "There is no CPU fallback" in r.stdout, r.stdout + r.stderr


def _tdg_host(pkg, linear):
    """A TimeDG with only its host-side constant builders (no device handle)."""
    from adjoint_ode_adaptivity_b200.tdg import TimeDG
    s = object.__new__(TimeDG)
    s.linear, s.quirks, s._cache, s._mesh_cache, s._h = linear, True, {}, {}, None
    return s


@pytest.mark.parametrize("linear", [False, True])
def test_tdg_padded_blocks_reproduce_mixed_orders(pkg, linear):
    """Mixed per-element orders (matlab/MAIN.m:21,141) are run by the device on blocks padded to
    the mesh maxima (layout: csrc/dgadj_tdg.cu).  NumPy emulation of that padded arithmetic --
    same blocks, same element recurrences -- against the oracle running each element at its
    own order: the padding must be inert."""
    from oracle import tdg as otdg
    rng = np.random.default_rng(5)
    times = np.array([0.0, 0.3, 0.8, 1.1, 2.0])
    Ks, Ns = 4, np.array([2, 1, 3, 1])
    y0 = rng.uniform(-3, 3, 6)
    s = _tdg_host(pkg, linear)
    consts, nodes, NP, nq = s.march_constants(Ns, times)
    blk = NP * NP + 2 * nq * NP + nq + 2
    assert consts.size == blk * Ks and NP == 4
    t1r, y1r, itsr = otdg.dg_march(Ns, Ks, times, y0, linear=linear)
    y = np.zeros((y0.size, Ks, NP))
    for b in range(y0.size):
        uR = y0[b]
        for k in range(Ks):
            c = consts[k * blk:(k + 1) * blk]
            A = c[:NP * NP].reshape(NP, NP)
            Iq = c[NP * NP:NP * NP + nq * NP].reshape(nq, NP)
            Phi = c[NP * NP + nq * NP:NP * NP + 2 * nq * NP].reshape(nq, NP)
            w = c[NP * NP + 2 * nq * NP:NP * NP + 2 * nq * NP + nq]
            hk, npk = c[-2], int(c[-1])
            assert npk == Ns[k] + 1
            F = np.zeros(NP); F[0] = uR
            if linear:
                U = np.linalg.solve(A, F)
            else:
                U = np.where(np.arange(NP) < npk, uR, 0.0)
                it, err = 0, 1.0
                while it <= 500 and err > 1e-7:
                    ur = Iq @ U
                    R = A @ U + hk / 2 * Phi.T @ (w * np.sin(ur)) + F
                    J = A + hk / 2 * Phi.T @ ((w * np.cos(ur))[:, None] * Phi)
                    dU = np.linalg.solve(J, R)
                    U = U - dU
                    err = np.linalg.norm(dU)
                    it += 1
                assert it == itsr[k][b]
            assert np.all(U[npk:] == 0.0)
            uR = U[npk - 1]
            y[b, k] = U
            np.testing.assert_allclose(U[:npk], y1r[k][b], rtol=1e-11, atol=1e-13)
    # adjoint blocks
    consts, nodes2, NPP, nq = s.adjoint_constants(Ns + 1, nodes)
    NA = NPP + 1
    blk = 2 * NA * NA + NA + NA * NPP + nq * NPP + nq * NA + nq + 3
    assert consts.size == blk * Ks and NPP == NP
    _, vr, errr = otdg.adj_march(Ns + 1, Ks, times, y1r, t1r, linear=linear)
    for b in range(y0.size):
        vL = 0.0
        for k in range(Ks - 1, -1, -1):
            c = consts[k * blk:(k + 1) * blk]
            o = 0
            A0 = c[o:o + NA * NA].reshape(NA, NA); o += NA * NA
            f1 = c[o:o + NA]; o += NA
            A2 = c[o:o + NA * NA].reshape(NA, NA); o += NA * NA
            Ix = c[o:o + NA * NPP].reshape(NA, NPP); o += NA * NPP
            Iq = c[o:o + nq * NPP].reshape(nq, NPP); o += nq * NPP
            Phi = c[o:o + nq * NA].reshape(nq, NA); o += nq * NA
            w = c[o:o + nq]; o += nq
            hk, nak, lastprev = c[o], int(c[o + 1]), int(c[o + 2])
            assert nak == Ns[k] + 2 and (k == 0 or lastprev == Ns[k - 1])
            Uk = y[b, k]
            ur = Iq @ Uk
            Mv = hk / 2 * Phi.T @ ((w * np.cos(ur))[:, None] * Phi) if nq else 0.0
            Mt = hk / 2 * Phi.T @ (w * np.sin(ur)) if nq else 0.0
            F = f1.copy(); F[nak - 1] -= vL
            vk = np.linalg.solve(A0 - Mv, F)
            vL = vk[0]
            F0 = np.zeros(NA); F0[0] = 1.0 if k == 0 else y[b, k - 1, lastprev]
            e = vk @ (-(A2 @ (Ix @ Uk)) - Mt + F0)
            assert np.all(vk[nak:] == 0.0)
            np.testing.assert_allclose(vk[:nak], vr[k][b], rtol=1e-9, atol=1e-11)
            np.testing.assert_allclose(e, errr[b, k], rtol=1e-9, atol=1e-11)


# ---------------------------------------------------------------- property tests of the host logic
def test_shard_range_partitions_any_batch(pkg):
    from hypothesis import given, settings, strategies as st
    from adjoint_ode_adaptivity_b200.sharding import shard_range

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 10**6), st.integers(1, 64))
    def check(B, world):
        cuts = [shard_range(B, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))            # contiguous, in rank order
        sizes = [hi - lo for lo, hi in cuts]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    check()


def test_refine_rules_insert_one_midpoint(pkg):
    """matlab/MAIN.m:137-141 / Main_finite_difference.py:336-341 on the host side: exactly one new
    point, the midpoint of the argmax element (lowest index on ties), everything else untouched."""
    from hypothesis import given, settings, strategies as st
    from adjoint_ode_adaptivity_b200.fd import refine_mesh
    from adjoint_ode_adaptivity_b200.tdg import refine

    @settings(max_examples=100, deadline=None)
    @given(st.lists(st.floats(0.01, 1.0), min_size=1, max_size=30), st.data())
    def check(widths, data):
        times = np.concatenate(([0.0], np.cumsum(widths)))
        K = len(widths)
        err = np.array(data.draw(st.lists(st.floats(-5, 5), min_size=K, max_size=K)))
        t2, Ns2, ref_i = refine(times, np.ones(K, dtype=int), err, 1)
        assert ref_i == int(np.argmax(np.abs(err))) and t2.size == times.size + 1 and Ns2.size == K + 1
        assert t2[ref_i + 1] == 0.5 * (times[ref_i] + times[ref_i + 1])
        assert np.array_equal(np.delete(t2, ref_i + 1), times) and np.all(np.diff(t2) > 0)
        assert np.array_equal(refine_mesh(times, ref_i), t2)
    check()


def test_fd_callable_probing(pkg):
    """The reference passes callables (python/Main_finite_difference.py:34,54,79); the host recognises the problem
    functions of its __main__ block by probing them, and refuses anything else."""
    from adjoint_ode_adaptivity_b200 import fd
    sin_rule = lambda u, dt, n: u[n - 1] + np.sin(u[n - 1]) * dt[n - 1]                 # :131-132
    lin_rule = lambda u, dt, n: (1 + dt[n - 1]) * u[n - 1]                             # :112-113
    assert fd._identify_ode(updateRule=sin_rule) == "sin" and fd._identify_ode(updateRule=lin_rule) == "linear"
    assert fd._identify_ode(getJF=lambda u, dt: np.diag(1 + np.cos(u[:-1]) * dt, -1)) == "sin"     # :138-140
    assert fd._identify_ode(getJF=lambda u, dt: np.diag(1 + dt, -1)) == "linear"                   # :118-119
    with pytest.raises(ValueError):
        fd._identify_ode(updateRule=sin_rule, getJF=lambda u, dt: np.diag(1 + dt, -1))
    with pytest.raises(NotImplementedError):
        fd._identify_ode(updateRule=lambda u, dt, n: u[n - 1] * (1 - dt[n - 1]))
    assert fd._identify_functional(lambda dt, u: np.concatenate((2 * u[:-1] * dt, 0), axis=None)) == "int_u2"   # :225-227
    assert fd._identify_functional(lambda dt, u=None: np.concatenate((dt, 0), axis=None)) == "int_u"          # :153-155

    def getK_uN(dt, u=None):                                                                                  # :162-165
        k = np.zeros_like(dt)
        k[-1] = 1
        return np.concatenate((k, 0), axis=None)
    assert fd._identify_functional(getK_uN) == "u_N"
    with pytest.raises(NotImplementedError):
        fd._identify_functional(lambda dt, u: np.concatenate((u[:-1] ** 3 * dt, 0), axis=None))
    assert np.allclose(fd.interpU(None, np.array([1.0, 1.0]), np.array([0.0, 1.0, 3.0])), [0, .25, .5, .75, 1, 1.5, 2, 2.5, 3])
    m = pkg.matlab_names
    v3 = np.array([[1.0, -1.0, 2.0], [2.0, -3.0, -1.0], [0.5, -2.0, 1.0]])
    assert m.minmod(v3).tolist() == [0.5, -1.0, 0.0]
    with pytest.raises(RuntimeError):
        m.AdvecRHS1D(np.zeros((3, 4)), 0.0, 1.0) if m.G.advec is None else (_ for _ in ()).throw(RuntimeError())
