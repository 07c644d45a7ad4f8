"""Smoothed-aggregation AMG setup on the host -- ``smoothed_aggregation_solver`` for scalar (CSR) and vector
(BSR, several near-nullspace candidates) problems.

Mirror of pyamg/aggregation/aggregation.py:26-431 for its default pipeline: symmetric strength (for BSR on the
Frobenius norms of the blocks, strength.py:327-345) -> standard aggregation of the node graph -> candidates
improved by 4 symmetric (block) Gauss-Seidel sweeps on A B = 0 (finest level) -> tentative prolongator T by
orthonormalising the candidates over every aggregate (``fit_candidates``, tentative.py:11-152 /
smoothed_aggregation.h:484-600: modified Gram-Schmidt, K1 unknowns per node x K2 candidates) ->
P = (I - omega/rho(D^-1 A) D^-1 A) T -> R = P^T -> A_c = R A P -> ``MultilevelSolver`` + ``change_smoothers``.
Native pieces: csrc/host_setup.cpp.

Setup, i.e. NOT the accelerated path (see classical.py): it lets BASELINE configs[1] (Poisson 2000^2, SA +
weighted Jacobi) and configs[4] (linear elasticity, BSR, block Jacobi) be synthesised where the reference is
not installed.  The spectral-radius estimate is a seeded Arnoldi here (the reference's is randomly started), so
omega -- hence P -- agrees with a reference run only to the accuracy of that estimate unless the same ``rho``
values are injected (``rho=[...]``; what tests/test_setup.py does).  Other strength/aggregation/smoothing
choices, nonsymmetric problems: NotImplementedError.
"""
import numpy as np
from scipy import sparse

from . import _host as H
from .classical import _csr32
from .multilevel import MultilevelSolver
from .relaxation.smoothing import change_smoothers
from .util import approximate_spectral_radius, galerkin, get_diagonal

__all__ = ["smoothed_aggregation_solver", "symmetric_strength_pattern", "standard_aggregation",
           "fit_candidates", "jacobi_prolongation_smoother"]


def _bsr32(A):
    A = A if (sparse.issparse(A) and A.format == "bsr") else sparse.bsr_array(A)
    A = sparse.bsr_array((np.ascontiguousarray(A.data, dtype=np.float64), A.indices.astype(np.int32),
                          A.indptr.astype(np.int32)), shape=A.shape, blocksize=A.blocksize)
    return A


def _is_block(A):
    return sparse.issparse(A) and A.format == "bsr" and A.blocksize != (1, 1)


def symmetric_strength_pattern(A, theta=0.0):
    """CSR pattern (values 1) of the symmetric strength-of-connection graph of A, diagonal included.  BSR input:
    the graph of the NODES, measured on the Frobenius norms of the blocks (strength.py:327-345)."""
    if theta < 0:
        raise ValueError("expected a positive theta")
    if _is_block(A):
        R, C = A.blocksize
        if R != C:
            raise ValueError("matrix must have square blocks")
        shape = (A.shape[0] // R, A.shape[1] // C)
        if theta == 0:
            return sparse.csr_array((np.ones(len(A.indices)), A.indices.astype(np.int32), A.indptr.astype(np.int32)),
                                    shape=shape)
        norms = np.sqrt((A.data * A.data).reshape(-1, R * C).sum(axis=1))
        A = sparse.csr_array((norms, A.indices, A.indptr), shape=shape)
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
    """Tentative prolongator T and coarse candidates R with T R = B on the aggregated nodes and orthonormal
    columns per aggregate (tentative.py:11-152).  B is (K1 * n_nodes, K2): K1 unknowns per node, K2 candidates.
    Per aggregate the K2 columns of the stacked candidate rows are orthonormalised by modified Gram-Schmidt in
    column order; a column whose norm after orthogonalisation is not above tol x (its norm before) is zeroed
    (smoothed_aggregation.h:516-600).  T is CSR for K1 = K2 = 1, else BSR with (K1, K2) blocks."""
    B = np.asarray(B, dtype=np.float64)
    if B.ndim == 1:
        B = B.reshape(-1, 1)
    AggOp = _csr32(AggOp)
    n, na = AggOp.shape
    if B.shape[0] % n != 0:
        raise ValueError(f"Dimensions of AggOp {AggOp.shape} and B {B.shape} are incompatible")
    K1, K2 = B.shape[0] // n, B.shape[1]
    member = np.diff(AggOp.indptr) > 0
    nodes = np.nonzero(member)[0]
    agg_of = AggOp.indices                                        # aggregate of nodes[k]
    order = np.argsort(agg_of, kind="stable")                     # aggregate-major, nodes ascending inside
    Q = B.reshape(n, K1, K2)[nodes[order]].copy()                 # (members, K1, K2)
    gid = np.repeat(agg_of[order], K1)                            # aggregate of every stacked row
    Qr = Q.reshape(-1, K2)                                        # view: stacked rows x candidates
    R = np.zeros((na, K2, K2))
    colsum = lambda w: np.bincount(gid, weights=w, minlength=na)
    for bj in range(K2):
        norm0 = np.sqrt(colsum(Qr[:, bj] * Qr[:, bj]))
        for bi in range(bj):
            d = colsum(Qr[:, bj] * Qr[:, bi])
            Qr[:, bj] -= d[gid] * Qr[:, bi]
            R[:, bi, bj] = d
        norm1 = np.sqrt(colsum(Qr[:, bj] * Qr[:, bj]))
        keep = norm1 > tol * norm0
        scale = np.zeros(na)
        scale[keep] = 1.0 / norm1[keep]
        R[keep, bj, bj] = norm1[keep]
        Qr[:, bj] *= scale[gid]
    data = np.empty_like(Q)
    data[order] = Q                                               # back to node order
    indptr = np.concatenate([[0], np.cumsum(member)]).astype(np.int32)
    if K1 == 1 and K2 == 1:
        T = sparse.csr_array((data.reshape(-1), agg_of.copy(), indptr), shape=(n, na))
    else:
        T = sparse.bsr_array((data, agg_of.copy(), indptr), shape=(n * K1, na * K2), blocksize=(K1, K2))
    return T, R.reshape(-1, K2)


def jacobi_prolongation_smoother(S, T, omega=4.0 / 3.0, degree=1, rho=None):
    """P = (I - omega/rho(D^-1 S) D^-1 S)^degree T   (diagonal weighting, smooth.py 'diagonal'); block operators
    keep their block structure."""
    D_inv = get_diagonal(S, inv=True)
    S_given = S
    if _is_block(S):
        S = _bsr32(S)
        Rb = S.blocksize[0]
        rows = np.repeat(np.arange(S.shape[0] // Rb), np.diff(S.indptr))
        scale = D_inv.reshape(-1, Rb)[rows]                        # (blocks, Rb): one factor per block row
        D_inv_S = sparse.bsr_array((S.data * scale[:, :, None], S.indices, S.indptr), shape=S.shape,
                                   blocksize=S.blocksize)
    else:
        S = _csr32(S)
        D_inv_S = sparse.dia_array((D_inv, 0), shape=S.shape) @ S
    if rho is None:
        if not _is_block(S):
            # the same seeded estimate setup_jacobi asks for later: computed once, cached on the operator
            from .relaxation.smoothing import rho_D_inv_A
            rho = rho_D_inv_A(S_given) if sparse.issparse(S_given) and S_given.format == "csr" else rho_D_inv_A(S)
        else:
            rho = approximate_spectral_radius(D_inv_S)
    D_inv_S = (omega / rho) * D_inv_S
    P = T
    for _ in range(degree):
        P = P - D_inv_S @ P
    if sparse.issparse(T) and T.format == "bsr":
        return _bsr32(P.tobsr(blocksize=T.blocksize))
    return _csr32(P)


def _improve_candidates(A, B, fn, kw):
    """B <- (relaxation on A x = 0 started from every column of B)  (aggregation.py:359-367)."""
    if fn not in ("gauss_seidel", "block_gauss_seidel") or kw.get("sweep", "forward") not in ("symmetric", "forward"):
        raise NotImplementedError("host SA setup improves candidates with (block_)gauss_seidel only")
    its = int(kw.get("iterations", 1))
    sym = 1 if kw.get("sweep", "forward") == "symmetric" else 0
    B = np.array(B, dtype=np.float64, order="F")                   # columns contiguous
    zero = np.zeros(A.shape[0])
    bs = A.blocksize[0] if (fn == "block_gauss_seidel" and _is_block(A)) else 1
    if bs == 1:
        Ac = _csr32(A)
        for k in range(B.shape[1]):
            H.lib().amgb_setup_gauss_seidel(Ac.shape[0], H.ip(Ac.indptr), H.ip(Ac.indices), H.dp(Ac.data),
                                            H.dp(B[:, k]), H.dp(zero), its, sym)
    else:
        Ab = _bsr32(A)
        from .util import get_block_diag
        Dinv = get_block_diag(Ab, blocksize=bs, inv_flag=True)
        data = np.ascontiguousarray(Ab.data)
        for k in range(B.shape[1]):
            H.lib().amgb_setup_block_gauss_seidel(Ab.shape[0] // bs, bs, H.ip(Ab.indptr), H.ip(Ab.indices),
                                                  H.dp(data.reshape(-1)), H.dp(Dinv.reshape(-1)), H.dp(B[:, k]),
                                                  H.dp(zero), its, sym)
    return np.ascontiguousarray(B)


def smoothed_aggregation_solver(A, B=None, BH=None, symmetry="hermitian", strength="symmetric", aggregate="standard",
                                smooth=("jacobi", {"omega": 4.0 / 3.0}),
                                presmoother=("block_gauss_seidel", {"sweep": "symmetric"}),
                                postsmoother=("block_gauss_seidel", {"sweep": "symmetric"}),
                                improve_candidates=(("block_gauss_seidel", {"sweep": "symmetric", "iterations": 4}),
                                                    None),
                                max_levels=10, max_coarse=10, diagonal_dominance=False, keep=False, rho=None,
                                **kwargs):
    """Create a multilevel solver using classical-style smoothed aggregation -- the reference's signature and
    defaults (aggregation.py:26-40), symmetric block Gauss-Seidel smoothing included (the engine runs it as
    dependency waves of the block graph).  ``BH`` (left near-nullspace of nonsymmetric problems) and
    ``diagonal_dominance`` belong to setup variants the host setup does not offer: they raise instead of being
    ignored.  ``rho``: optional list of known spectral radii rho(D^-1 A) per level (skips the estimates)."""
    if BH is not None:
        raise NotImplementedError("host SA setup: BH (nonsymmetric problems) -- build with the reference and adopt "
                                  "the hierarchy with MultilevelSolver.from_pyamg")
    if diagonal_dominance:
        raise NotImplementedError("host SA setup: diagonal_dominance -- build with the reference and adopt the "
                                  "hierarchy with MultilevelSolver.from_pyamg")
    def unpack(v):
        return (v[0], v[1]) if isinstance(v, tuple) else (v, {})

    def levelize(spec):
        if isinstance(spec, list):
            return [spec[min(i, len(spec) - 1)] for i in range(max_levels)]
        if isinstance(spec, tuple) and len(spec) == 2 and not isinstance(spec[1], dict):
            return [spec[min(i, 1)] for i in range(max_levels)]      # (first, rest) form of improve_candidates
        return [spec] * max_levels

    if symmetry not in ("hermitian", "symmetric"):
        raise NotImplementedError("host SA setup: symmetric problems only")
    A = _bsr32(A) if _is_block(A) else _csr32(A)
    if A.shape[0] != A.shape[1]:
        raise ValueError("expected square matrix")
    bs0 = A.blocksize[0] if _is_block(A) else 1
    if B is None:
        B = np.kron(np.ones((A.shape[0] // bs0, 1)), np.eye(bs0))      # aggregation.py:222-225
    B = np.asarray(B, dtype=np.float64)
    if B.ndim == 1:
        B = B.reshape(-1, 1)
    if B.shape[0] != A.shape[0]:
        raise ValueError("The shape of near null-space modes B is incorrect")
    strength, aggregate, smooth = levelize(strength), levelize(aggregate), levelize(smooth)
    improve_candidates = levelize(improve_candidates)
    rho = list(rho) if rho is not None else []

    def nodes_of(M):
        return M.shape[0] // (M.blocksize[0] if _is_block(M) else 1)

    levels = [MultilevelSolver.Level()]
    levels[-1].A, levels[-1].B = A, B
    while len(levels) < max_levels and nodes_of(levels[-1].A) > max_coarse:
        k = len(levels) - 1
        A, B = levels[-1].A, levels[-1].B
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
            B = _improve_candidates(A, B, fn, kw)
            levels[-1].B = B
        T, Bc = fit_candidates(AggOp, B)
        fn, kw = unpack(smooth[k])
        if fn == "jacobi":
            if set(kw) - {"omega", "degree"}:
                raise NotImplementedError("host SA setup: jacobi prolongation smoothing with omega/degree only")
            P = jacobi_prolongation_smoother(A, T, rho=rho[k] if k < len(rho) else None, **kw)
        elif fn is None:
            P = _bsr32(T) if T.format == "bsr" else _csr32(T)
        else:
            raise NotImplementedError("host SA setup offers smooth=('jacobi', {...}) or None only")
        if P.format == "bsr":
            R = _bsr32(P.T.tobsr(blocksize=P.blocksize[::-1]))
            Ac = (R @ A @ P)
            Ac = _bsr32(Ac.tobsr(blocksize=(P.blocksize[1], P.blocksize[1])))
        else:
            R = _csr32(P.T.tocsr())
            Ac = _csr32(galerkin(R, A, P) if (A.format == "csr" and R.format == "csr") else R @ A @ P)
        if keep:
            levels[-1].C, levels[-1].AggOp, levels[-1].Cnodes, levels[-1].T = C, AggOp, Cnodes, T
        levels[-1].P, levels[-1].R = P, R
        levels.append(MultilevelSolver.Level())
        levels[-1].A = Ac
        levels[-1].B = Bc
    ml = MultilevelSolver(levels, **kwargs)
    change_smoothers(ml, presmoother, postsmoother)
    return ml
