"""CPU oracle for the dgadj hot path -- TEST INFRASTRUCTURE ONLY.

NumPy fp64 restatement of the reference's algorithms (wglao/Adjoint-ODE-Adaptivity,
`utils/*.m`, `utils/One_code.mlx`, `matlab/{fem_setup,dg_march,adj_march,adj_rec,err_contribution}.m`,
`python/Main_finite_difference.py`).  Every function cites the reference file:line it
follows.  Nothing in the product package may import this; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs do,
and there only as the checker / reported baseline.

Parity status:
  * operators + forward LSERK4 advection: PINNED by the goldens embedded in
    `utils/One_code.mlx` (4 significant digits) -> tests/golden/mlx_one_code.json
  * finite-difference path: PINNED at fp64 by fixtures generated from the reference's
    own `python/Main_finite_difference.py` -> tests/golden/fd_reference.json
  * time-DG dg_march/adj_march: weakly pinned (init_nonlin.png, 3 digits)
  * adj_rec.m (disabled in the reference) and err_contribution.m (never called, needs the
    Symbolic Toolbox): PARITY UNPINNED -- no reference output exists; checked against the exact
    adjoint / the exact functional of the linear model problem
  * PDE discrete adjoint + indicator, Burgers RHS, periodic BC, upwind flux:
    PARITY UNPINNED -- the reference has no such code; the oracle is build-defined
    and validated by dot-product / finite-difference / effectivity identities.
"""
