"""Classical (Ruge-Stueben) AMG setup on the host -- ``ruge_stuben_solver``.

Mirror of pyamg/classical/classical.py:20-203 for its DEFAULT pipeline: classical strength
(theta = 0.25) -> first-pass RS C/F splitting -> modified classical interpolation -> R = P^T ->
Galerkin product A_c = R A P (SciPy SpGEMM, as in the reference :201) -> ``MultilevelSolver`` +
``change_smoothers``.  The algorithmic kernels are csrc/host_setup.cpp.  This is setup, i.e. NOT
the accelerated path: it exists to synthesise the BASELINE hierarchies where the reference is not
installed (GPU box); hierarchies built by the reference itself are adopted with
``MultilevelSolver.from_pyamg``.  Other strength / splitting / interpolation choices of the
reference are not offered here (NotImplementedError).
"""
import numpy as np
from scipy import sparse

from . import _host as H
from .multilevel import MultilevelSolver
from .util import galerkin
from .relaxation.smoothing import change_smoothers

__all__ = ["ruge_stuben_solver", "classical_strength_of_connection", "RS", "classical_interpolation"]


def _csr32(A):
    A = sparse.csr_array(A)
    if A.dtype != np.float64:
        A = A.astype(np.float64)
    A.indptr = np.ascontiguousarray(A.indptr, dtype=np.int32)
    A.indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    A.data = np.ascontiguousarray(A.data, dtype=np.float64)
    return A


def classical_strength_of_connection(A, theta=0.25, return_index=False):
    """Strength matrix: |a_ij| >= theta * max_{k!=i} |a_ik| (diagonal kept), scaled row-wise to max 1."""
    if theta < 0 or theta > 1:
        raise ValueError("expected theta in [0,1]")
    A = _csr32(A)
    n = A.shape[0]
    Sp = np.empty(n + 1, dtype=np.int32)
    Sj = np.empty(A.nnz, dtype=np.int32)
    Sx = np.empty(A.nnz, dtype=np.float64)
    Sidx = np.empty(A.nnz, dtype=np.int32)
    nnz = H.lib().amgb_setup_classical_strength(n, H.ip(A.indptr), H.ip(A.indices), H.dp(A.data),
                                                float(theta), H.ip(Sp), H.ip(Sj), H.dp(Sx), H.ip(Sidx))
    S = sparse.csr_array((Sx[:nnz], Sj[:nnz], Sp), shape=(n, n))
    if return_index:
        return S, Sidx[:nnz]
    return S


def RS(S):
    """First-pass Ruge-Stueben C/F splitting of the strength graph S (1 = C point, 0 = F point)."""
    S = _csr32(S)
    n = S.shape[0]
    # drop the diagonal (split.RS: remove_diagonal), keep the row order
    rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(S.indptr))
    off = rows != S.indices
    Sj = np.ascontiguousarray(S.indices[off])
    Sp = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows[off], minlength=n), out=Sp[1:])
    So = sparse.csr_array((np.ones(len(Sj)), Sj, Sp), shape=(n, n))
    So.sort_indices()            # the reference's remove_diagonal round-trips through COO -> sorted rows;
    Sj = np.ascontiguousarray(So.indices, dtype=np.int32)   # the visiting order decides ties
    T = _csr32(So.T.tocsr())
    splitting = np.empty(n, dtype=np.int32)
    H.lib().amgb_setup_rs_splitting(n, H.ip(Sp), H.ip(Sj), H.ip(T.indptr), H.ip(T.indices), H.ip(splitting))
    return splitting


def classical_interpolation(A, S, Sidx, splitting):
    """Modified classical (distance-1) interpolation on the strength pattern S of A."""
    A = _csr32(A)
    S = _csr32(S)
    n = A.shape[0]
    splitting = np.ascontiguousarray(splitting, dtype=np.int32)
    keep = np.empty(S.nnz, dtype=np.uint8)
    Pp = np.empty(n + 1, dtype=np.int32)
    L = H.lib()
    L.amgb_setup_classical_interp_count(n, H.ip(S.indptr), H.ip(S.indices), H.ip(splitting), H.u8p(keep), H.ip(Pp))
    Pj = np.empty(Pp[-1], dtype=np.int32)
    Px = np.empty(Pp[-1], dtype=np.float64)
    Sv = np.ascontiguousarray(A.data[Sidx])
    L.amgb_setup_classical_interp_fill(n, H.ip(A.indptr), H.ip(A.indices), H.dp(A.data), H.ip(S.indptr),
                                       H.ip(S.indices), H.dp(Sv), H.u8p(keep), H.ip(splitting), H.ip(Pp),
                                       H.ip(Pj), H.dp(Px))
    return sparse.csr_array((Px, Pj, Pp), shape=(n, int(splitting.sum())))


def ruge_stuben_solver(A, strength=("classical", {"theta": 0.25}), CF=("RS", {"second_pass": False}),
                       interpolation="classical",
                       presmoother=("gauss_seidel", {"sweep": "symmetric"}),
                       postsmoother=("gauss_seidel", {"sweep": "symmetric"}),
                       max_levels=30, max_coarse=10, keep=False, **kwargs):
    """Create a multilevel solver using classical (Ruge-Stueben) AMG -- same signature and defaults
    as pyamg.ruge_stuben_solver (classical.py:20-26)."""
    def unpack(v):
        return (v[0], v[1]) if isinstance(v, tuple) else (v, {})

    sfn, skw = unpack(strength)
    cfn, ckw = unpack(CF)
    ifn, ikw = unpack(interpolation)
    if sfn != "classical" or set(skw) - {"theta"}:
        raise NotImplementedError("host setup offers strength=('classical', {'theta': t}) only")
    if cfn != "RS" or ckw.get("second_pass", False):
        raise NotImplementedError("host setup offers CF=('RS', {'second_pass': False}) only")
    if ifn != "classical" or ikw:
        raise NotImplementedError("host setup offers interpolation='classical' (modified) only")
    if not sparse.issparse(A) or A.format != "csr":
        A = sparse.csr_array(A)
    A = _csr32(A)
    if A.shape[0] != A.shape[1]:
        raise ValueError("expected square matrix")

    levels = [MultilevelSolver.Level()]
    levels[-1].A = A
    while len(levels) < max_levels and levels[-1].A.shape[0] > max_coarse:
        A = levels[-1].A
        S, Sidx = classical_strength_of_connection(A, return_index=True, **skw)
        splitting = RS(S)
        nc = int(splitting.sum())
        if nc == 0 or nc == len(splitting):
            break                                           # classical.py:174-177
        P = classical_interpolation(A, S, Sidx, splitting)
        R = _csr32(P.T.tocsr())                             # classical.py:189
        if keep:
            levels[-1].C = S
        levels[-1].splitting = splitting.astype(bool)
        levels[-1].P = P
        levels[-1].R = R
        levels.append(MultilevelSolver.Level())
        levels[-1].A = _csr32(galerkin(R, A, P))            # classical.py:201 (AMGB_GPU_RAP=1: on the GPU)
    ml = MultilevelSolver(levels, **kwargs)
    change_smoothers(ml, presmoother, postsmoother)
    return ml
