"""dgadj -- B200-native batched 1-D nodal-DG forward / adjoint march with adjoint-weighted error
indicators: a from-scratch drop-in for the hot path of wglao/Adjoint-ODE-Adaptivity.

The directory is named `adjoint-ode-adaptivity_b200` (not an importable identifier); load it
with `dgadj_loader.load_package()` from the repo root, which registers it as the module
`adjoint_ode_adaptivity_b200`.
"""
from . import _lib
from ._lib import DgadjError
from .galerkin import BaseGalerkin1D
from .solver import AdvecDG1D
from .fd import FDAdjoint, refine_mesh
from .tdg import TimeDG
from .burgers import BurgersDG1D, decode_limiter_record
from .adapt import adapt_fd, adapt_tdg, adapt_advec, adapt_fd_per_trajectory, adapt_tdg_per_trajectory
from . import matlab_names
from .sharding import (shard_range, allreduce_indicators, batch_mean_refine, gather_indicators,
                       allreduce_indicator_blocks, combine_blocks, REDUCE_BLOCK)

__all__ = ["AdvecDG1D", "BaseGalerkin1D", "DgadjError", "FDAdjoint", "refine_mesh", "TimeDG", "BurgersDG1D", "decode_limiter_record", "adapt_fd", "adapt_tdg", "adapt_advec", "adapt_fd_per_trajectory", "adapt_tdg_per_trajectory", "shard_range", "allreduce_indicators",
           "batch_mean_refine", "gather_indicators", "allreduce_indicator_blocks", "combine_blocks", "REDUCE_BLOCK", "_lib"]
