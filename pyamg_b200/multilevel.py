"""MultilevelSolver: the reference's hierarchy container and cycle API, executed on a B200.

Mirror of pyamg/multilevel.py.  The object API is kept (``levels`` list of ``Level`` structs with
``A, P, R, presmoother, postsmoother``; ``solve`` :398-582 with the same keyword arguments, return
values and stop test; ``aspreconditioner`` :355-396; ``psolve``; ``change_solve_matrix``;
``__repr__`` and the three complexity measures :184-318), but ``solve`` no longer recurses through
NumPy/SciPy: the hierarchy is uploaded once to HBM and the cycle (pre-smooth, residual,
restriction, coarse solve, prolongation + correction, post-smooth) runs as sm_100a CUDA kernels
captured in a CUDA graph (csrc/engine.cu).  Host arrays in, a new host array out.

Build it from
  * a reference solver:  ``MultilevelSolver.from_pyamg(pyamg.ruge_stuben_solver(A))``
    (setup stays on the reference's CPU path; closures are parsed for omega / Dinv / row lists), or
  * a list of ``Level`` objects exactly as the reference's constructors do, followed by
    ``pyamg_b200.relaxation.smoothing.change_smoothers``.

No CPU fallback: without libpyamg_b200.so and a CUDA device ``solve`` raises.
"""
import ctypes
import os
from warnings import warn

import numpy as np
import scipy.linalg
import scipy.sparse.linalg as sla
from scipy.sparse.linalg import LinearOperator

from . import _engine as E
from .relaxation import smoothing

__all__ = ["MultilevelSolver", "coarse_grid_solver"]


def coarse_solver_spec(cs):
    """The ``coarse_solver=`` argument a coarse solver object was built from: ``name`` or ``(name, kwargs)``.

    The reference's ``GenericSolver.name()`` returns only the method name (multilevel.py:820-823); its keyword
    arguments -- e.g. the sweep count of a relaxation coarse solver -- live in the closure of the ``solve`` function
    its ``__call__`` uses (:700-790).  They are read from there, so that a reference-built hierarchy is adopted with
    the coarse solver it really has."""
    import ast
    own = getattr(cs, "spec", None)
    if own is not None:
        return own
    name = cs.name() if hasattr(cs, "name") else "'pinv'"
    try:
        name = ast.literal_eval(name)
    except (ValueError, SyntaxError) as exc:
        raise NotImplementedError(f"coarse solver {name} is not on the GPU hot path") from exc
    kwargs = {}
    try:
        call = type(cs).__call__
        cells = dict(zip(call.__code__.co_freevars, [c.cell_contents for c in call.__closure__ or ()]))
        solve = cells.get("solve")
        if solve is not None and solve.__closure__:
            kwargs = dict(zip(solve.__code__.co_freevars, [c.cell_contents for c in solve.__closure__])).get("kwargs", {})
    except (AttributeError, TypeError, ValueError):
        kwargs = {}
    if isinstance(name, tuple):
        return name
    return (name, dict(kwargs)) if kwargs else name


# relaxation methods usable as the coarsest-level solver (multilevel.py:764-766) that the engine runs
_RELAXATION_COARSE = ("gauss_seidel", "jacobi", "block_gauss_seidel", "block_jacobi", "richardson", "sor", "chebyshev")


def coarse_grid_solver(solver):
    """Coarse-level solver descriptor (pyamg/multilevel.py:665-826).

    The engine applies the coarsest solve as a dense matrix-vector product with a matrix computed
    once on the CPU: the pseudo-inverse for 'pinv'/'pinv2' (what the reference caches, :717-721),
    the plain inverse for the direct methods 'lu'/'cholesky'/'splu' (same solve, different
    rounding).  Relaxation methods (:764-781: x = 0, then `iterations` sweeps, default 10) run as the engine's
    smoother kernels on the coarsest level.  Krylov coarse solvers are outside the accelerated path.
    """
    def unpack_arg(v):
        if isinstance(v, tuple):
            return v[0], v[1]
        return v, {}

    name, kwargs = unpack_arg(solver)
    if name in ("pinv", "pinv2"):
        def dense(A):
            return scipy.linalg.pinv(A.toarray(), **kwargs)
    elif name in ("lu", "cholesky", "splu"):
        def dense(A):
            return scipy.linalg.inv(A.toarray())
    elif name is None:
        def dense(A):
            return np.zeros(A.shape)
    elif name in _RELAXATION_COARSE:
        dense = None                                  # x = 0, then `iterations` (default 10) sweeps on the GPU
        kwargs = dict(kwargs)
        kwargs.setdefault("iterations", 10)           # multilevel.py:768-769
    elif isinstance(name, str) or callable(name):
        raise NotImplementedError(f"coarse solver {name!r} is not on the GPU hot path; use 'pinv' "
                                  "(the reference's default), 'lu', 'cholesky', 'splu', a relaxation method "
                                  f"({', '.join(_RELAXATION_COARSE)}) or None")
    else:
        raise ValueError(f"unknown solver: {name}")

    relax_spec = (name, kwargs) if dense is None else None
    relax_name = name                                 # (`name` itself is shadowed inside the class body below)

    class GenericSolver:
        """Holds the cached dense operator (or the relaxation descriptor); applied on the GPU by the cycle engine."""
        relaxation = relax_spec
        spec = solver                                 # what this object was built from (see coarse_solver_spec)
        if relax_spec is not None:
            P = None                                  # no dense operator: the solve is `iterations` GPU sweeps

        def smoother(self, A):
            """The closure the reference builds for a relaxation coarse solver (multilevel.py:773-776)."""
            lvl = MultilevelSolver.Level()
            lvl.A = A
            return smoothing._setup_call(relax_name)(lvl, **kwargs)

        def dense_operator(self, A):
            if not hasattr(self, "P"):
                self.P = np.ascontiguousarray(dense(A), dtype=np.float64)
            return self.P

        def __repr__(self):
            return "coarse_grid_solver(" + repr(solver) + ")"

        @classmethod
        def name(cls):
            return repr(solver)

    return GenericSolver()


def _cg_host(A, b, x0=None, tol=1e-5, maxiter=None, M=None, callback=None, residuals=None):
    """pyamg.krylov.cg (pyamg/krylov/_cg.py:11-196, criteria 'rr') on the host -- the path taken when a
    callback needs host iterates; ``M`` is the GPU cycle as a LinearOperator.  Same steps as amgb_solve_cg."""
    b = np.ravel(np.asarray(b, dtype=np.float64))
    n = len(b)
    x = np.zeros(n) if x0 is None else np.array(np.ravel(x0), dtype=np.float64)
    if maxiter is None:
        maxiter = int(1.3 * n) + 2
    elif maxiter < 1:
        raise ValueError("Number of iterations must be positive")
    r = b - A @ x
    z = M @ r
    p = z.copy()
    rz = np.inner(r, z)
    normr = np.linalg.norm(r)
    if residuals is not None:
        residuals[:] = [normr]
    normb = np.linalg.norm(b)
    if normb == 0.0:
        normb = 1.0
    rtol = tol * normb
    if normr < rtol:
        return x, 0
    it = 0
    while True:
        Ap = A @ p
        rz_old = rz
        pAp = np.inner(Ap, p)
        if pAp < 0.0:
            warn("\nIndefinite matrix detected in CG, aborting\n")
            return x, -1
        alpha = rz / pAp
        x += alpha * p
        if np.mod(it, 8) and it > 0:
            r -= alpha * Ap
        else:
            r = b - A @ x
        z = M @ r
        rz = np.inner(r, z)
        if rz < 0.0:
            warn("\nIndefinite preconditioner detected in CG, aborting\n")
            return x, -1
        beta = rz / rz_old
        p *= beta
        p += z
        it += 1
        normr = np.linalg.norm(r)
        if residuals is not None:
            residuals.append(normr)
        if callback is not None:
            callback(x)
        if normr < rtol:
            return x, 0
        if it == maxiter:
            return x, it


class MultilevelSolver:
    """Stores a multigrid hierarchy and runs the multigrid cycle on the GPU.

    Parameters and attributes as pyamg.multilevel.MultilevelSolver (multilevel.py:17-182).
    """

    class Level:
        """One level of the hierarchy: a struct with A (+ P, R, presmoother, postsmoother)."""

        def __init__(self):
            self.A = None

    def __init__(self, levels, coarse_solver="pinv", device=0, stream=None):
        self.symmetric_smoothing = False
        self.levels = levels
        self.coarse_solver = coarse_grid_solver(coarse_solver)
        self.device = device
        self._stream = stream
        self._h = None
        for level in levels[:-1]:
            if not hasattr(level, "R"):
                level.R = level.P.T.conjugate()     # multilevel.py:180-182

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_pyamg(cls, ml, device=0, stream=None):
        """Adopt a hierarchy built by the reference (any ``pyamg`` constructor): same Level
        objects, same smoother closures, same coarse solver kind (and its cached pinv if any)."""
        spec = coarse_solver_spec(ml.coarse_solver)
        new = cls(list(ml.levels), coarse_solver=spec, device=device, stream=stream)
        new.symmetric_smoothing = getattr(ml, "symmetric_smoothing", False)
        cached = getattr(ml.coarse_solver, "P", None)
        if cached is not None:
            new.coarse_solver.P = np.ascontiguousarray(cached, dtype=np.float64)
        return new

    def _invalidate(self):
        """Drop the device copy (levels or smoothers changed); re-uploaded on the next solve."""
        if self._h is not None:
            E.lib().amgb_hierarchy_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def upload(self):
        """Upload the hierarchy to HBM (idempotent). Returns the device footprint in bytes."""
        if self._h is not None:
            return int(E.lib().amgb_hierarchy_device_bytes(self._h))
        E.require_gpu()
        L = E.lib()
        h = ctypes.c_void_p()
        E.check(L.amgb_hierarchy_create(int(self.device), ctypes.byref(h)))
        try:
            nlv = len(self.levels)
            for k, lvl in enumerate(self.levels):
                keep = []
                A = E.as_matrix(lvl.A, keep)
                if k < nlv - 1:
                    P = E.as_matrix(lvl.P, keep)
                    R = E.as_matrix(lvl.R, keep)
                    pre = smoothing.describe(getattr(lvl, "presmoother", None), lvl.A, keep)
                    post = smoothing.describe(getattr(lvl, "postsmoother", None), lvl.A, keep)
                    E.check(L.amgb_hierarchy_add_level(h, ctypes.byref(A), ctypes.byref(P),
                                                       ctypes.byref(R), ctypes.byref(pre),
                                                       ctypes.byref(post)))
                else:
                    E.check(L.amgb_hierarchy_add_level(h, ctypes.byref(A), None, None, None, None))
            Ac = self.levels[-1].A
            if Ac.nnz == 0:     # GenericSolver.__call__, multilevel.py:801-803
                E.check(L.amgb_hierarchy_set_coarse_pinv(h, Ac.shape[0], None, 1))
            elif getattr(self.coarse_solver, "relaxation", None) is not None:
                keep = []
                S = smoothing.describe(self.coarse_solver.smoother(Ac), Ac, keep)
                E.check(L.amgb_hierarchy_set_coarse_relaxation(h, ctypes.byref(S)))
            else:
                Pd = self.coarse_solver.dense_operator(Ac)
                E.check(L.amgb_hierarchy_set_coarse_pinv(h, Ac.shape[0], E.f64p(Pd.reshape(-1)), 0))
            E.check(L.amgb_hierarchy_finalize(h, ctypes.c_void_p(self._stream or 0)))
        except Exception:
            L.amgb_hierarchy_destroy(h)
            raise
        self._h = h
        return int(L.amgb_hierarchy_device_bytes(h))

    @property
    def handle(self):
        self.upload()
        return self._h

    def last_launches(self):
        """CUDA kernels launched by the most recent solve (graph nodes counted individually)."""
        return int(E.lib().amgb_hierarchy_last_launches(self._h)) if self._h is not None else 0

    def profile_cycle(self, cycle="V", max_records=200000):
        """Time every operator launch of one (un-graphed) cycle with CUDA events.

        Returns a float array (n, 6): level, op (0 spmv/restrict, 1 residual, 2 prolong+add, 3 jacobi,
        4 gs wave, 5 block jacobi, 6 cluster tail, 7 resident GS, 8 indexed Jacobi, 9 block GS wave), rows, nnz,
        algorithmic bytes, milliseconds."""
        rec = np.empty(max_records * 6, dtype=np.float64)
        nrec = ctypes.c_int32(0)
        E.check(E.lib().amgb_profile_cycle(self.handle, E.CYCLES[str(cycle).upper()], E.f64p(rec),
                                           int(max_records), ctypes.byref(nrec)))
        return rec[:nrec.value * 6].reshape(-1, 6).copy()

    # ------------------------------------------------------------------ reporting (host only)
    def __repr__(self):
        """Basic statistics of the hierarchy (multilevel.py:184-209)."""
        output = "MultilevelSolver\n"
        output += f"Number of Levels:     {len(self.levels)}\n"
        output += f"Operator Complexity:  {self.operator_complexity():6.3f}\n"
        output += f"Grid Complexity:      {self.grid_complexity():6.3f}\n"
        output += f"Coarse Solver:        {self.coarse_solver.name()}\n"
        total_nnz = sum(level.A.nnz for level in self.levels)
        output += "  level   unknowns     nonzeros\n"
        for n, level in enumerate(self.levels):
            A = level.A
            ratio = 100 * A.nnz / total_nnz
            output += f"{n:>6} {A.shape[1]:>11} {A.nnz:>12} [{ratio:2.2f}%]\n"
        return output

    def cycle_complexity(self, cycle="V"):
        """Nonzeros touched by one cycle relative to the fine level (multilevel.py:211-283): every visit
        of a non-coarsest level costs 2 nnz (pre + post smoothing), a coarse solve nnz of the coarsest level;
        V visits each level once, W twice per parent visit, F = one F-visit plus one V-visit of the next level."""
        cycle = str(cycle).upper()
        if cycle not in ("V", "W", "F", "AMLI"):
            raise TypeError(f"Unrecognized cycle type ({cycle})")
        nnz = [level.A.nnz for level in self.levels]
        if len(nnz) == 1:
            return 1.0
        # bottom-up: cost of one visit of level l under each recursion pattern
        v = w = f = 2 * nnz[-2] + nnz[-1]
        for l in range(len(nnz) - 3, -1, -1):
            v, w, f = 2 * nnz[l] + v, 2 * nnz[l] + 2 * w, 2 * nnz[l] + f + v
        flops = {"V": v, "W": w, "AMLI": w, "F": f}[cycle]
        return float(flops) / float(nnz[0])

    def operator_complexity(self):
        return sum(level.A.nnz for level in self.levels) / float(self.levels[0].A.nnz)

    def grid_complexity(self):
        return sum(level.A.shape[0] for level in self.levels) / float(self.levels[0].A.shape[0])

    # ------------------------------------------------------------------ solve-phase API
    def change_solve_matrix(self, A):
        """Change the fine-level matrix and rebuild its smoothers (multilevel.py:320-337)."""
        self.levels[0].A = A
        smoothing.rebuild_smoother(self.levels[0])
        self._invalidate()

    def _solve_cg_device(self, b, x0, tol, maxiter, cycle, residuals, return_info, method="cg"):
        """solve(accel='cg' | 'bicgstab') on the GPU: pyamg's CG (pyamg/krylov/_cg.py) or BiCGStab (_bicgstab.py)
        preconditioned by one cycle."""
        b = np.asarray(b)
        n = self.levels[0].A.shape[0]
        if b.size != n or (x0 is not None and np.asarray(x0).size != n):
            raise ValueError("b / x0 have invalid dimensions")
        if maxiter is None:
            maxiter = int(1.3 * n) + 2 if method == "cg" else n + 5     # _cg.py:92-93, _bicgstab.py:90-91
        elif maxiter < 1:
            raise ValueError("Number of iterations must be positive")
        bh = np.ascontiguousarray(np.ravel(b), dtype=np.float64)
        xh = np.zeros(n) if x0 is None else np.array(np.ravel(x0), dtype=np.float64)
        res = np.empty(int(maxiter) + 1, dtype=np.float64)
        nres, info = ctypes.c_int32(0), ctypes.c_int32(0)
        fn = E.lib().amgb_solve_cg if method == "cg" else E.lib().amgb_solve_bicgstab
        E.check(fn(self.handle, bh.ctypes.data, xh.ctypes.data, float(tol), int(maxiter),
                   E.CYCLES[cycle], E.FLAG_X0_ZERO if x0 is None else 0, E.f64p(res),
                   ctypes.byref(nres), ctypes.byref(info)))
        if info.value == -1:
            warn("\nIndefinite matrix or preconditioner detected in CG, aborting\n")
        if residuals is not None:
            residuals[:] = list(res[:nres.value])
        xout = xh.ravel()          # the reference's Krylov solvers hand back the ravelled iterate
        return (xout, info.value) if return_info else xout

    def _solve_gmres_device(self, b, x0, tol, maxiter, cycle, residuals, return_info, flexible):
        """solve(accel='gmres' | 'fgmres') on the GPU: pyamg's Householder GMRES / flexible GMRES
        (pyamg/krylov/_gmres_householder.py, _fgmres.py) preconditioned by one cycle."""
        b = np.asarray(b)
        n = self.levels[0].A.shape[0]
        if b.size != n or (x0 is not None and np.asarray(x0).size != n):
            raise ValueError("b / x0 have invalid dimensions")
        if maxiter is not None and maxiter > n:
            warn("Setting maxiter to maximum allowed, n.")            # _gmres_householder.py:146-148
        max_inner = min(n, 40) if maxiter is None else min(int(maxiter), n)
        bh = np.ascontiguousarray(np.ravel(b), dtype=np.float64)
        xh = np.zeros(n) if x0 is None else np.array(np.ravel(x0), dtype=np.float64)
        res = np.empty(max_inner + 3, dtype=np.float64)
        nres, info = ctypes.c_int32(0), ctypes.c_int32(0)
        flags = (E.FLAG_X0_ZERO if x0 is None else 0) | (E.FLAG_FLEXIBLE if flexible else 0)
        E.check(E.lib().amgb_solve_gmres(self.handle, bh.ctypes.data, xh.ctypes.data, float(tol), 0,
                                         0 if maxiter is None else int(maxiter), E.CYCLES[cycle], flags,
                                         E.f64p(res), len(res), ctypes.byref(nres), ctypes.byref(info)))
        if residuals is not None:
            residuals[:] = list(res[:min(nres.value, len(res))])
        xout = xh.ravel()          # the reference's Krylov solvers hand back the ravelled iterate
        return (xout, info.value) if return_info else xout

    def psolve(self, b):
        """Legacy interface: one iteration (multilevel.py:339-353)."""
        return self.solve(b, maxiter=1)

    def aspreconditioner(self, cycle="V"):
        """LinearOperator applying one cycle from x0 = 0 (multilevel.py:355-396)."""
        shape = self.levels[0].A.shape
        dtype = self.levels[0].A.dtype

        def matvec(b):
            return self.solve(b, maxiter=1, cycle=cycle, tol=1e-12)

        M = LinearOperator(shape, matvec, dtype=dtype)
        M._amgb_solver, M._amgb_cycle = self, str(cycle).upper()     # lets pyamg_b200.krylov keep the solve resident
        return M

    def solve(self, b, x0=None, tol=1e-5, maxiter=100, cycle="V", accel=None, callback=None,
              residuals=None, cycles_per_level=1, return_info=False, out=None):
        """Execute multigrid cycling on the GPU (pyamg/multilevel.py:398-582; same arguments).

        ``out`` (extension, optional): a C-contiguous float64 array of n entries that receives the solution
        and is returned (reshaped like b) instead of a freshly allocated array -- with a page-locked buffer
        (``pyamg_b200.pinned_empty``) this removes the page-fault cost of a new 100+ MB result per call."""
        b = np.asarray(b)
        if out is not None:
            if accel is not None or callback is not None:
                raise ValueError("out= is supported for plain cycling only")
            if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous
                    and out.size == b.size):
                raise ValueError("out must be a C-contiguous float64 array with as many entries as b")
        if x0 is None:
            # the engine starts from zero on the device (AMGB_FLAG_X0_ZERO): the host buffer is output only
            x = out if out is not None else (np.empty_like(b, dtype=np.float64) if accel is None and callback is None
                                             and not np.iscomplexobj(b) else np.zeros_like(b))
        else:
            x = np.array(x0)    # copy (:467)
            if out is not None:
                np.copyto(out.reshape(x.shape), x)
                x = out

        A = self.levels[0].A
        cycle = str(cycle).upper()

        if cycle not in E.CYCLES:
            raise TypeError(f"Unrecognized cycle type ({cycle})")      # :658
        # AMLI cycles require a hermitian matrix (multilevel.py:472-476)
        if cycle == "AMLI" and hasattr(A, "symmetry") and A.symmetry != "hermitian":
            raise ValueError("AMLI cycles require symmetry to be hermitian")

        if accel is not None:
            if accel != "fgmres" and cycle == "AMLI":                  # multilevel.py:487-490
                raise ValueError("AMLI cycles require acceleration (accel) to be fgmres, or no acceleration")
            # Check for symmetric smoothing scheme when using CG (multilevel.py:481-485)
            if (accel == "cg") and (not self.symmetric_smoothing):
                warn("Incompatible non-symmetric multigrid preconditioner "
                     "detected, due to presmoother/postsmoother combination. "
                     "CG requires SPD preconditioner, not just SPD matrix.")
            if accel == "cg" and callback is None and not np.iscomplexobj(b):
                # pyamg.krylov.cg (what the reference resolves 'cg' to, multilevel.py:495-499) with every
                # vector resident in HBM: amgb_solve_cg
                return self._solve_cg_device(b, x0, tol, maxiter, cycle, residuals, return_info)
            if accel == "bicgstab" and callback is None and not np.iscomplexobj(b) \
                    and self.levels[0].A.shape[0] > 1:
                # pyamg.krylov.bicgstab (right-preconditioned, criteria 'rr') resident: amgb_solve_bicgstab
                return self._solve_cg_device(b, x0, tol, maxiter, cycle, residuals, return_info, method="bicgstab")
            if accel in ("gmres", "fgmres") and callback is None and not np.iscomplexobj(b) \
                    and self.levels[0].A.shape[0] > 1:
                # pyamg.krylov.gmres (Householder) / pyamg.krylov.fgmres, what the reference resolves these
                # strings to (multilevel.py:495-499), with every long vector resident in HBM: amgb_solve_gmres
                return self._solve_gmres_device(b, x0, tol, maxiter, cycle, residuals, return_info,
                                                flexible=(accel == "fgmres"))
            if accel == "cg":
                accel = _cg_host           # same algorithm on the host (callback wants host iterates)
            elif isinstance(accel, str):
                # the reference looks the name up in pyamg.krylov first, then in scipy.sparse.linalg
                # (multilevel.py:495-499); its host solvers are used when the reference is importable, with the GPU
                # cycle as M (one host <-> device round trip per application)
                try:
                    from pyamg import krylov as ref_krylov
                except ImportError:
                    ref_krylov = None
                if ref_krylov is not None and hasattr(ref_krylov, accel):
                    accel = getattr(ref_krylov, accel)
                elif hasattr(sla, accel):
                    accel = getattr(sla, accel)
                else:
                    raise NotImplementedError(f"accel='{accel}': GPU-resident accelerators are 'cg', 'gmres' and "
                                              "'fgmres'; other names need pyamg.krylov or scipy.sparse.linalg to "
                                              "provide a host solver of that name")
            M = self.aspreconditioner(cycle=cycle)
            try:  # PyAMG style interface which has a residuals parameter (multilevel.py:503-508)
                x, info = accel(A, b, x0=x0, tol=tol, maxiter=maxiter, M=M, callback=callback,
                                residuals=residuals)
                if return_info:
                    return x, info
                return x
            except TypeError:
                # scipy.sparse.linalg style interface (multilevel.py:509-535)
                if residuals is not None:
                    residuals[:] = [np.linalg.norm(b - A @ x)]

                    def callback_wrapper(xk):
                        if np.isscalar(xk):
                            residuals.append(xk)
                        else:
                            residuals.append(np.linalg.norm(b - A @ xk))
                        if callback is not None:
                            callback(xk)
                else:
                    callback_wrapper = callback
                x, info = accel(A, b, x0=x0, maxiter=maxiter, M=M, callback=callback_wrapper,
                                rtol=tol, atol=0)
                if return_info:
                    return x, info
                return x

        if np.iscomplexobj(b) or np.iscomplexobj(x):
            raise NotImplementedError("complex systems are outside the fp64 hot path")
        n = A.shape[0]
        if b.size != n or x.size != n:
            raise ValueError("b / x0 have invalid dimensions")
        if maxiter < 1:
            raise ValueError("maxiter must be at least 1")
        out_shape = b.shape
        bh = np.ascontiguousarray(np.ravel(b), dtype=np.float64)        # :551-554
        xh = np.ascontiguousarray(np.ravel(x), dtype=np.float64)

        L = E.lib()
        h = self.handle
        cyc = E.CYCLES[cycle]
        info = ctypes.c_int32(0)
        nres = ctypes.c_int32(0)

        if callback is None:
            res = np.empty(maxiter + 1, dtype=np.float64)
            flags = E.FLAG_X0_ZERO if x0 is None else 0       # x0 = 0: no host->device copy of the guess
            E.check(L.amgb_solve_ex(h, bh.ctypes.data, xh.ctypes.data, float(tol), int(maxiter), cyc,
                                    int(cycles_per_level), flags, E.f64p(res), ctypes.byref(nres),
                                    ctypes.byref(info)))
            if residuals is not None:
                residuals[:] = list(res[:nres.value])
            status = info.value
        else:
            # callback(x) wants the host iterate after every cycle: one engine call per cycle
            res = np.empty(2, dtype=np.float64)
            it = 0
            normb = np.linalg.norm(bh)
            if normb == 0.0:
                normb = 1.0
            while True:
                E.check(L.amgb_solve(h, bh.ctypes.data, xh.ctypes.data, 0.0, 1, cyc,
                                     int(cycles_per_level), E.f64p(res), ctypes.byref(nres),
                                     ctypes.byref(info)))
                if it == 0 and residuals is not None:
                    residuals[:] = [res[0]]
                it += 1
                if residuals is not None:
                    residuals.append(res[1])
                callback(xh)
                if res[1] < tol * normb:
                    status = 0
                    break
                if it == maxiter:
                    status = it
                    break

        xout = xh.ravel()          # b and x are ravelled before cycling (multilevel.py:553-554): (n,) out, even for (n,1) in
        if return_info:
            return xout, status
        return xout
