"""Smoothed-aggregation AMG setup on the host for SCALAR problems -- ``smoothed_aggregation_solver``.

Mirror of pyamg/aggregation/aggregation.py:26-431 for its default pipeline with one near-nullspace
candidate: symmetric strength (theta = 0) -> standard aggregation -> candidates improved by 4 symmetric
Gauss-Seidel sweeps on A B = 0 (finest level) -> tentative prolongator T by normalising B over every
aggregate (``fit_candidates`` with K1 = K2 = 1) -> P = (I - omega/rho(D^-1 A) D^-1 A) T -> R = P^T ->
A_c = R A P -> ``MultilevelSolver`` + ``change_smoothers``.  Native pieces: csrc/host_setup.cpp.

Setup, i.e. NOT the accelerated path (see classical.py): it lets BASELINE configs[1] (Poisson 2000^2, SA +
weighted Jacobi) be synthesised where the reference is not installed.  The spectral-radius estimate is a
seeded Arnoldi here (the reference's is randomly started), so omega -- hence P -- agrees with a reference
run only to the accuracy of that estimate unless the same ``rho`` values are injected (``rho=[...]``; what
tests/test_setup.py does).  Block (BSR) problems, other strength/aggregation/smoothing choices:
NotImplementedError.
"""
import numpy as np
from scipy import sparse

from . import _host as H
from .classical import _csr32
from .multilevel import MultilevelSolver
from .relaxation.smoothing import change_smoothers
from .util import approximate_spectral_radius, get_diagonal

__all__ = ["smoothed_aggregation_solver", "symmetric_strength_pattern", "standard_aggregation",
           "fit_candidates", "jacobi_prolongation_smoother"]


def symmetric_strength_pattern(A, theta=0.0):
    """CSR pattern (values 1) of the symmetric strength-of-connection graph of A, diagonal included."""
    if theta < 0:
        raise ValueError("expected a positive theta")
    A = _csr32(A)
    n = A.shape[0]
    Sp = np.empty(n + 1, dtype=np.int32)
    Sj = np.empty(A.nnz, dtype=np.int32)
    nnz = H.lib().amgb_setup_symmetric_strength(n, H.ip(A.indptr), H.ip(A.indices), H.dp(A.data), float(theta),
                                                H.ip(Sp), H.ip(Sj))
    return sparse.csr_array((np.ones(nnz), Sj[:nnz].copy(), Sp), shape=(n, n))


def standard_aggregation(C):
    """(AggOp, roots): AggOp[i, a] = 1 iff node i belongs to aggregate a (isolated nodes: empty rows)."""
    C = _csr32(C)
    n = C.shape[0]
    agg = np.empty(n, dtype=np.int32)
    roots = np.empty(n, dtype=np.int32)
    na = H.lib().amgb_setup_standard_aggregation(n, H.ip(C.indptr), H.ip(C.indices), H.ip(agg), H.ip(roots))
    if na == 0:
        return sparse.csr_array((n, 1), dtype=np.int32), np.array([], dtype=np.int32)
    member = agg >= 0
    indptr = np.concatenate([[0], np.cumsum(member)]).astype(np.int32)
    AggOp = sparse.csr_array((np.ones(int(member.sum()), dtype=np.int32), agg[member], indptr), shape=(n, na))
    return AggOp, roots[:na].copy()


def fit_candidates(AggOp, B, tol=1e-10):
    """Tentative prolongator for ONE candidate: T[i, a] = B_i / ||B restricted to a||, coarse candidate
    R_a = that norm (columns whose norm is below tol are zeroed)."""
    B = np.asarray(B, dtype=np.float64).reshape(-1)
    AggOp = _csr32(AggOp)
    n, na = AggOp.shape
    if len(B) != n:
        raise NotImplementedError("fit_candidates: one candidate and one unknown per node only")
    rows = np.repeat(np.arange(n), np.diff(AggOp.indptr))
    aid = AggOp.indices
    norms = np.sqrt(np.bincount(aid, weights=B[rows] ** 2, minlength=na))
    keep = norms > tol * norms          # the reference's threshold is relative to the column's own norm
    scale = np.zeros(na)
    scale[keep] = 1.0 / norms[keep]
    R = np.where(keep, norms, 0.0).reshape(-1, 1)
    T = sparse.csr_array((B[rows] * scale[aid], aid.copy(), AggOp.indptr.copy()), shape=(n, na))
    return T, R


def jacobi_prolongation_smoother(S, T, omega=4.0 / 3.0, degree=1, rho=None):
    """P = (I - omega/rho(D^-1 S) D^-1 S)^degree T   (diagonal weighting)."""
    S = _csr32(S)
    D_inv = get_diagonal(S, inv=True)
    D_inv_S = sparse.dia_array((D_inv, 0), shape=S.shape) @ S
    if rho is None:
        rho = approximate_spectral_radius(D_inv_S)
    D_inv_S = (omega / rho) * D_inv_S
    P = T
    for _ in range(degree):
        P = P - D_inv_S @ P
    return _csr32(P)


def smoothed_aggregation_solver(A, B=None, symmetry="hermitian", strength="symmetric", aggregate="standard",
                                smooth=("jacobi", {"omega": 4.0 / 3.0}),
                                presmoother=("jacobi", {"omega": 4.0 / 3.0}),
                                postsmoother=("jacobi", {"omega": 4.0 / 3.0}),
                                improve_candidates=(("block_gauss_seidel", {"sweep": "symmetric", "iterations": 4}),
                                                    None),
                                max_levels=10, max_coarse=10, keep=False, rho=None, **kwargs):
    """Create a multilevel solver using classical-style smoothed aggregation -- the reference's signature
    (aggregation.py:26-40) for scalar problems; the default smoothers are weighted Jacobi here because the
    reference's default (lexicographic block Gauss-Seidel) is not a throughput smoother on a GPU (it is
    supported by the engine, as dependency waves, when requested)."""
    def unpack(v):
        return (v[0], v[1]) if isinstance(v, tuple) else (v, {})

    def levelize(spec):
        if isinstance(spec, list):
            return [spec[min(i, len(spec) - 1)] for i in range(max_levels)]
        if isinstance(spec, tuple) and len(spec) == 2 and not isinstance(spec[1], dict):
            return [spec[min(i, 1)] for i in range(max_levels)]      # (first, rest) form of improve_candidates
        return [spec] * max_levels

    if sparse.issparse(A) and A.format == "bsr" and A.blocksize != (1, 1):
        raise NotImplementedError("host SA setup handles scalar problems (one unknown per node) only")
    if symmetry not in ("hermitian", "symmetric"):
        raise NotImplementedError("host SA setup: symmetric problems only")
    A = _csr32(A)
    if A.shape[0] != A.shape[1]:
        raise ValueError("expected square matrix")
    B = np.ones(A.shape[0]) if B is None else np.asarray(B, dtype=np.float64).reshape(-1)
    if len(B) != A.shape[0]:
        raise NotImplementedError("host SA setup: one near-nullspace candidate only")
    strength, aggregate, smooth = levelize(strength), levelize(aggregate), levelize(smooth)
    improve_candidates = levelize(improve_candidates)
    rho = list(rho) if rho is not None else []

    levels = [MultilevelSolver.Level()]
    levels[-1].A, levels[-1].B = A, B.reshape(-1, 1)
    while len(levels) < max_levels and levels[-1].A.shape[0] > max_coarse:
        k = len(levels) - 1
        A, B = levels[-1].A, levels[-1].B.reshape(-1)
        fn, kw = unpack(strength[k])
        if fn != "symmetric" or set(kw) - {"theta"}:
            raise NotImplementedError("host SA setup offers strength=('symmetric', {'theta': t}) only")
        C = symmetric_strength_pattern(A, **kw)
        fn, kw = unpack(aggregate[k])
        if fn != "standard" or kw:
            raise NotImplementedError("host SA setup offers aggregate='standard' only")
        AggOp, Cnodes = standard_aggregation(C)
        fn, kw = unpack(improve_candidates[k])
        if fn is not None:
            if fn not in ("gauss_seidel", "block_gauss_seidel") or kw.get("sweep", "forward") not in ("symmetric", "forward"):
                raise NotImplementedError("host SA setup improves candidates with (block_)gauss_seidel only")
            B = B.copy()
            H.lib().amgb_setup_gauss_seidel(A.shape[0], H.ip(A.indptr), H.ip(A.indices), H.dp(A.data), H.dp(B),
                                            H.dp(np.zeros(A.shape[0])), int(kw.get("iterations", 1)),
                                            1 if kw.get("sweep", "forward") == "symmetric" else 0)
            levels[-1].B = B.reshape(-1, 1)
        T, Bc = fit_candidates(AggOp, B)
        fn, kw = unpack(smooth[k])
        if fn == "jacobi":
            if set(kw) - {"omega", "degree"}:
                raise NotImplementedError("host SA setup: jacobi prolongation smoothing with omega/degree only")
            P = jacobi_prolongation_smoother(A, T, rho=rho[k] if k < len(rho) else None, **kw)
        elif fn is None:
            P = _csr32(T)
        else:
            raise NotImplementedError("host SA setup offers smooth=('jacobi', {...}) or None only")
        R = _csr32(P.T.tocsr())
        if keep:
            levels[-1].C, levels[-1].AggOp, levels[-1].Cnodes, levels[-1].T = C, AggOp, Cnodes, T
        levels[-1].P, levels[-1].R = P, R
        levels.append(MultilevelSolver.Level())
        levels[-1].A = _csr32(R @ A @ P)
        levels[-1].B = Bc
    ml = MultilevelSolver(levels, **kwargs)
    change_smoothers(ml, presmoother, postsmoother)
    return ml
