"""Drive the UNMODIFIED reference (baseline/_ref: /root/reference's pyamg compiled in place by oracle/build.py)
on a hierarchy this repo holds.  TEST INFRASTRUCTURE ONLY -- bench.py's `--impl reference` / `cpu_baseline` legs and
tests/ use it; the product never imports it.

``to_reference(ml)`` hands the levels' operators to ``pyamg.MultilevelSolver(levels, coarse_solver)``
(pyamg/multilevel.py:165-182) and installs on every level exactly what the reference's own ``change_smoothers``
would store there: ``functools.partial(pyamg.relaxation.relaxation.<name>, **kwargs)`` (pyamg/relaxation/
smoothing.py:494-579).  From there on every call is the reference's stock code path: ``ml_ref.solve`` :398-582 ->
``__solve`` :584-662 -> ``relaxation.<name>`` -> ``amg_core.<name>`` (relaxation.h) and SciPy's ``csr_matvec``.
"""
from functools import partial, update_wrapper

import numpy as np

from .build import import_reference

# smoother names (as oracle.smoother_spec reports them) the reference's relaxation module provides under the same name
_SAME_NAME = ("jacobi", "gauss_seidel", "sor", "gauss_seidel_indexed", "block_jacobi", "block_gauss_seidel",
              "jacobi_indexed", "cf_jacobi", "fc_jacobi", "polynomial")


def _reference_smoother(pyamg, spec):
    if spec is None:
        def none(A, x, b):
            return None
        return none
    name, kw = spec
    if name not in _SAME_NAME:
        raise NotImplementedError(f"reference adapter: smoother '{name}'")
    fn = getattr(pyamg.relaxation.relaxation, name)
    kw = dict(kw)
    if name == "gauss_seidel_indexed":
        kw["indices"] = np.ascontiguousarray(kw["indices"], dtype=np.int32)
    sm = partial(fn, **kw)
    update_wrapper(sm, fn)
    return sm


def to_reference(ml, coarse_solver="pinv"):
    """pyamg.MultilevelSolver over the operators and smoother parameters of `ml` (a pyamg_b200.MultilevelSolver)."""
    from . import hierarchy_spec
    pyamg = import_reference()
    levels = []
    for d in hierarchy_spec(ml):
        lvl = pyamg.MultilevelSolver.Level()
        lvl.A = d["A"]
        if "P" in d:
            lvl.P, lvl.R = d["P"], d["R"]
        levels.append((lvl, d))
    ref = pyamg.MultilevelSolver([lv for lv, _ in levels], coarse_solver=coarse_solver)
    for lvl, d in levels:
        if "P" in d:
            lvl.presmoother = _reference_smoother(pyamg, d["pre"])
            lvl.postsmoother = _reference_smoother(pyamg, d["post"])
    return ref
