"""Small cases of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dgadj_loader
pkg = dgadj_loader.load_package()
dev = "cuda"
rng = np.random.default_rng(0)
for (N, K, bc, ept) in [(3, 8, "periodic", 0), (2, 7, "inflow", 0), (8, 64, "periodic", 0), (4, 16, "inflow", 1)]:
    s = pkg.AdvecDG1D(N, K, domain=(0.0, 1.0), alpha=0.0, bc=bc)
    if ept:
        s.set_tuning(elems_per_thread=ept)
    u0 = torch.tensor(rng.standard_normal((5, N + 1, K)), device=dev)
    out = s.fwd_adj(u0, 1.0, 1e-3, 6, want_lam0=True)
    uT, ck = s.forward_checkpointed(u0, 1.0, 1e-3, 6)
    o2 = s.adjoint(uT, ck, 1.0, 1e-3, 6)
    assert torch.equal(o2["eta"], out["eta"])
    s.forward(u0, 1.0, 1e-3, 4, history=True)
    s.rank(out["eta"], 2)
    s.reduce_indicators(out["eta"], out["J"])
    s.rhs(u0, 0.1, 1.0)
    h = s.fwd_adj(u0.cpu().numpy(), 1.0, 1e-3, 6)
f = pkg.FDAdjoint()
f.solve(torch.tensor(rng.uniform(-3, 3, 100), device=dev), np.diff(np.linspace(0, 2, 9)))
t = pkg.TimeDG()
times = np.linspace(0, 2, 5); Ns = np.ones(4, dtype=int)
t1, y1, its = t.dg_march(Ns, 4, times, torch.tensor(rng.uniform(0.5, 2, 40), device=dev))
t.adj_march(Ns + 1, 4, times, y1, t1)
b = pkg.BurgersDG1D(3, 33, bc="free")
b.forward(torch.tensor(rng.standard_normal((3, 4, 33)), device=dev), 1e-3, 5, history=True, checkpoints=True)
torch.cuda.synchronize()
print("sanitize_case ok")
