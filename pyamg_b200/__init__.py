"""pyamg_b200 -- a B200-native engine for the algebraic-multigrid SOLVE PHASE behind pyamg's API.

Scope: ``MultilevelSolver.solve`` / the V-cycle and its smoothers + SpMVs run as
hand-written sm_100a CUDA kernels (csrc/); hierarchy setup stays on the reference's CPU path (or
on this package's small host-side setup used to synthesise benchmark inputs).  There is no CPU
fallback for the solve phase.
"""
from .multilevel import MultilevelSolver, coarse_grid_solver
from . import relaxation
from .relaxation.smoothing import change_smoothers
from ._engine import pinned_empty, EngineError
from . import gallery
from . import krylov
from .classical import ruge_stuben_solver
from .aggregation import smoothed_aggregation_solver

__version__ = "0.1.0"
__all__ = ["MultilevelSolver", "coarse_grid_solver", "relaxation", "change_smoothers",
           "pinned_empty", "EngineError", "gallery", "krylov", "ruge_stuben_solver", "smoothed_aggregation_solver"]
