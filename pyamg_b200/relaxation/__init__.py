"""Relaxation (smoother) layer: GPU sweeps behind the reference's Python signatures.

Mirrors pyamg/relaxation/{relaxation,smoothing}.py for the smoothers on the hot path
(SURVEY.md 8(a) rows a5-a9, a11).
"""
from . import relaxation, smoothing
from .relaxation import (jacobi, gauss_seidel, gauss_seidel_indexed, block_jacobi, sor,
                         make_system)
from .smoothing import change_smoothers

__all__ = ["relaxation", "smoothing", "jacobi", "gauss_seidel", "gauss_seidel_indexed",
           "block_jacobi", "sor", "make_system", "change_smoothers"]
