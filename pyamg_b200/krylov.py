"""Krylov solvers with the call signatures of ``pyamg.krylov`` whose preconditioner is a pyamg_b200 cycle.

The usual way to use the reference's hierarchy as a preconditioner is

    x, info = pyamg.krylov.gmres(A, b, M=ml.aspreconditioner(cycle='V'), tol=1e-8, residuals=res)

which, with a GPU hierarchy behind ``M``, would bounce every Krylov vector host <-> device once per iteration.
The functions below take the same arguments; when ``M`` comes from ``pyamg_b200.MultilevelSolver.aspreconditioner``
and ``A`` is that hierarchy's fine-level operator they run the whole accelerated solve resident in HBM
(``amgb_solve_cg`` / ``amgb_solve_gmres``: pyamg/krylov/_cg.py, _gmres_householder.py, _fgmres.py restated, see
csrc/abi_krylov.cuh).  Anything else is outside the accelerated path and raises -- there is no CPU fallback here;
SciPy's / the reference's host solvers accept the same ``M`` unchanged.
"""
import numpy as np

__all__ = ["cg", "gmres", "fgmres", "bicgstab"]


def _resident(A, M, name):
    ml = getattr(M, "_amgb_solver", None)
    if ml is None:
        raise NotImplementedError(f"pyamg_b200.krylov.{name}: M must come from pyamg_b200.MultilevelSolver.aspreconditioner "
                                  "(host Krylov solvers accept that operator too)")
    A0 = ml.levels[0].A
    same = A is A0
    if not same and getattr(A, "shape", None) == A0.shape:
        try:
            same = (abs(A - A0)).nnz == 0
        except (TypeError, ValueError):
            same = False
    if not same:
        raise NotImplementedError(f"pyamg_b200.krylov.{name}: A must be the fine-level operator of the preconditioning "
                                  "hierarchy (ml.levels[0].A)")
    return ml, M._amgb_cycle


def _run(name, A, b, x0, tol, maxiter, M, callback, residuals, restart=None):
    if callback is not None:
        raise NotImplementedError(f"pyamg_b200.krylov.{name}: callbacks need host iterates; use ml.solve(accel=...)")
    if restart is not None:
        raise NotImplementedError(f"pyamg_b200.krylov.{name}: restarts are not exposed (amgb_solve_gmres has them)")
    ml, cycle = _resident(A, M, name)
    b = np.asarray(b)
    x, info = ml.solve(b, x0=x0, tol=tol, maxiter=maxiter, cycle=cycle, accel=name, residuals=residuals,
                       return_info=True)
    return x.reshape(b.shape), info


def cg(A, b, x0=None, tol=1e-5, criteria="rr", maxiter=None, M=None, callback=None, residuals=None):
    """pyamg.krylov.cg (krylov/_cg.py:11-196) with the cycle ``M``, resident on the GPU (criteria 'rr' only)."""
    if criteria != "rr":
        raise NotImplementedError("pyamg_b200.krylov.cg: stopping criteria 'rr' only")
    return _run("cg", A, b, x0, tol, maxiter, M, callback, residuals)


def gmres(A, b, x0=None, tol=1e-5, restart=None, maxiter=None, M=None, callback=None, residuals=None,
          orthog="householder", **kwargs):
    """pyamg.krylov.gmres (krylov/_gmres.py:11-127 -> _gmres_householder.py) with the cycle ``M``, resident."""
    if orthog != "householder":
        raise NotImplementedError("pyamg_b200.krylov.gmres: orthog='householder' (the reference's default) only")
    return _run("gmres", A, b, x0, tol, maxiter, M, callback, residuals, restart)


def fgmres(A, b, x0=None, tol=1e-5, restart=None, maxiter=None, M=None, callback=None, residuals=None, **kwargs):
    """pyamg.krylov.fgmres (krylov/_fgmres.py:17-345) with the cycle ``M``, resident."""
    return _run("fgmres", A, b, x0, tol, maxiter, M, callback, residuals, restart)


def bicgstab(A, b, x0=None, tol=1e-5, criteria="rr", maxiter=None, M=None, callback=None, residuals=None):
    """pyamg.krylov.bicgstab (krylov/_bicgstab.py:10-200) with the cycle ``M``, resident (criteria 'rr' only)."""
    if criteria != "rr":
        raise NotImplementedError("pyamg_b200.krylov.bicgstab: stopping criteria 'rr' only")
    return _run("bicgstab", A, b, x0, tol, maxiter, M, callback, residuals)
