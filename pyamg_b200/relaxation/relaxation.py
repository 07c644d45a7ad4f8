"""GPU relaxation sweeps with the call signatures of pyamg/relaxation/relaxation.py.

Each function validates its arguments exactly like the reference's ``make_system``
(relaxation.py:15-97: ValueError for non-arrays / wrong sizes / non-contiguous x, TypeError for
mixed dtypes), then hands HOST arrays to the reference-FFI-shaped C entry points of
libpyamg_b200.so (``amgb_host_*``), which upload, run the sm_100a kernels and write x back in
place.  They exist so the reference's smoother tests can run against the CUDA kernels unchanged;
inside a multigrid cycle the same kernels run on resident data (multilevel.MultilevelSolver).

There is no CPU path: without a GPU these raise.
"""
from warnings import warn

import numpy as np
from scipy import sparse

from .. import _engine as E

__all__ = ["make_system", "jacobi", "gauss_seidel", "gauss_seidel_indexed", "block_jacobi", "sor",
           "polynomial", "jacobi_indexed", "cf_jacobi", "fc_jacobi", "block_gauss_seidel", "cf_block_jacobi",
           "fc_block_jacobi", "jacobi_ne", "gauss_seidel_ne", "gauss_seidel_nr"]


def make_system(A, x, b, formats=None):
    """Return A,x,b suitable for relaxation or raise an exception (relaxation.py:15-97)."""
    if formats is None:
        pass
    elif formats == ["csr"]:
        if sparse.issparse(A) and A.format == "csr":
            pass
        elif sparse.issparse(A) and A.format == "bsr":
            A = A.tocsr()
        else:
            warn("implicit conversion to CSR", sparse.SparseEfficiencyWarning)
            A = sparse.csr_array(A)
    elif sparse.issparse(A) and A.format in formats:
        pass
    else:
        A = sparse.csr_array(A).asformat(formats[0])

    if not isinstance(x, np.ndarray):
        raise ValueError("expected numpy array for argument x")
    if not isinstance(b, np.ndarray):
        raise ValueError("expected numpy array for argument b")
    M, N = A.shape
    if M != N:
        raise ValueError("expected square matrix")
    if x.shape not in [(M,), (M, 1)]:
        raise ValueError("x has invalid dimensions")
    if b.shape not in [(M,), (M, 1)]:
        raise ValueError("b has invalid dimensions")
    if A.dtype != x.dtype or A.dtype != b.dtype:
        raise TypeError("arguments A, x, and b must have the same dtype")
    if not x.flags.carray:
        raise ValueError("x must be contiguous in memory")
    return A, np.ravel(x), np.ravel(b)


def _fp64(A):
    """The engine computes in fp64 only (BASELINE.json); the reference's f32/complex overloads
    (instantiate.yml:2-6) are outside the accelerated path and fail loudly."""
    if A.dtype != np.float64:
        raise NotImplementedError(f"pyamg_b200 relaxes fp64 systems only (got {A.dtype})")


def _csr_args(A):
    Ap = np.ascontiguousarray(A.indptr, dtype=np.int32)
    Aj = np.ascontiguousarray(A.indices, dtype=np.int32)
    Ax = np.ascontiguousarray(A.data, dtype=np.float64).reshape(-1)
    return Ap, Aj, Ax


def jacobi(A, x, b, iterations=1, omega=1.0):
    """Weighted Jacobi, in place on x (relaxation.py:349-420 -> relaxation.h:309-346 / :472-562)."""
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _fp64(A)
    n = A.shape[0]
    if n == 0:
        return
    E.require_gpu()
    L = E.lib()
    temp = np.empty_like(x)
    om = np.array([omega], dtype=np.float64)
    b = np.ascontiguousarray(b)
    Ap, Aj, Ax = _csr_args(A)
    if A.format == "csr":
        for _ in range(iterations):
            E.check(L.amgb_host_jacobi(E.i32p(Ap), len(Ap), E.i32p(Aj), len(Aj), E.f64p(Ax), len(Ax),
                                       E.f64p(x), n, E.f64p(b), n, E.f64p(temp), n,
                                       0, n, 1, E.f64p(om), 1))
    else:
        R, C = A.blocksize
        if R != C:
            raise ValueError("BSR blocks must be square")
        for _ in range(iterations):
            E.check(L.amgb_host_bsr_jacobi(E.i32p(Ap), len(Ap), E.i32p(Aj), len(Aj), E.f64p(Ax),
                                           len(Ax), E.f64p(x), n, E.f64p(b), n, E.f64p(temp), n,
                                           0, n // R, 1, R, E.f64p(om), 1))


def gauss_seidel(A, x, b, iterations=1, sweep="forward", omega=1.0):
    """Gauss-Seidel, in place on x (relaxation.py:265-346 -> relaxation.h:48-76; omega != 1 is SOR,
    relaxation.h:116-145 -- and, as in the reference, the symmetric sweep ignores omega).

    The sequential sweep is executed as dependency waves (rows of a wave are mutually
    independent), which reproduces the lexicographic result.
    """
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _fp64(A)
    if A.format == "bsr":
        # point-wise GS on BSR == GS on the CSR expansion (pinned by test_relaxation.py:224-249)
        R, C = A.blocksize
        if R != C:
            raise ValueError("BSR blocks must be square")
        A = A.tocsr()
        omega = 1.0      # reference quirk: the BSR branch (bsr_gauss_seidel, relaxation.py:343-346) has no omega
    n = A.shape[0]
    if sweep == "forward":
        rs = (0, n, 1)
    elif sweep == "backward":
        rs = (n - 1, -1, -1)
    elif sweep == "symmetric":
        for _ in range(iterations):
            gauss_seidel(A, x, b, iterations=1, sweep="forward")
            gauss_seidel(A, x, b, iterations=1, sweep="backward")
        return
    else:
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
    if n == 0:
        return
    E.require_gpu()
    L = E.lib()
    b = np.ascontiguousarray(b)
    Ap, Aj, Ax = _csr_args(A)
    for _ in range(iterations):
        if omega != 1.0:
            E.check(L.amgb_host_sor_gauss_seidel(E.i32p(Ap), len(Ap), E.i32p(Aj), len(Aj), E.f64p(Ax),
                                                 len(Ax), E.f64p(x), n, E.f64p(b), n, *rs, float(omega)))
        else:
            E.check(L.amgb_host_gauss_seidel(E.i32p(Ap), len(Ap), E.i32p(Aj), len(Aj), E.f64p(Ax),
                                             len(Ax), E.f64p(x), n, E.f64p(b), n, *rs))


def sor(A, x, b, omega, iterations=1, sweep="forward"):
    """SOR (relaxation.py:100-154): Gauss-Seidel with the in-sweep damping omega."""
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    for _ in range(iterations):
        gauss_seidel(A, x, b, iterations=1, sweep=sweep, omega=omega)


def gauss_seidel_indexed(A, x, b, indices, iterations=1, sweep="forward"):
    """Gauss-Seidel over an explicit row list (relaxation.py:662-731 -> relaxation.h:736-768);
    with rows sorted by colour this is the reference's multi-colour GS."""
    A, x, b = make_system(A, x, b, formats=["csr"])
    _fp64(A)
    indices = np.ascontiguousarray(np.asarray(indices, dtype="intc"), dtype=np.int32)
    m = len(indices)
    if sweep == "forward":
        rs = (0, m, 1)
    elif sweep == "backward":
        rs = (m - 1, -1, -1)
    elif sweep == "symmetric":
        for _ in range(iterations):
            gauss_seidel_indexed(A, x, b, indices, iterations=1, sweep="forward")
            gauss_seidel_indexed(A, x, b, indices, iterations=1, sweep="backward")
        return
    else:
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
    if m == 0:
        return
    E.require_gpu()
    L = E.lib()
    n = A.shape[0]
    b = np.ascontiguousarray(b)
    Ap, Aj, Ax = _csr_args(A)
    for _ in range(iterations):
        E.check(L.amgb_host_gauss_seidel_indexed(E.i32p(Ap), len(Ap), E.i32p(Aj), len(Aj), E.f64p(Ax),
                                                 len(Ax), E.f64p(x), n, E.f64p(b), n,
                                                 E.i32p(indices), m, *rs))


def block_jacobi(A, x, b, Dinv=None, blocksize=1, iterations=1, omega=1.0):
    """Block Jacobi, in place on x (relaxation.py:423-499 -> relaxation.h:1021-1090)."""
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _fp64(A)
    A = A.tobsr(blocksize=(blocksize, blocksize))
    if Dinv is None:
        from ..util import get_block_diag
        Dinv = get_block_diag(A, blocksize=blocksize, inv_flag=True)
    elif Dinv.shape[0] != int(A.shape[0] / blocksize):
        raise ValueError("Dinv and A have incompatible dimensions")
    elif (Dinv.shape[1] != blocksize) or (Dinv.shape[2] != blocksize):
        raise ValueError("Dinv and blocksize are incompatible")
    nb = int(A.shape[0] / blocksize)
    if nb <= 0:
        return
    E.require_gpu()
    L = E.lib()
    n = A.shape[0]
    temp = np.empty_like(x)
    om = np.array([omega], dtype=np.float64)
    b = np.ascontiguousarray(b)
    Ap, Aj, Ax = _csr_args(A)
    Tx = np.ascontiguousarray(Dinv, dtype=np.float64).reshape(-1)
    for _ in range(iterations):
        E.check(L.amgb_host_block_jacobi(E.i32p(Ap), len(Ap), E.i32p(Aj), len(Aj), E.f64p(Ax), len(Ax),
                                         E.f64p(x), n, E.f64p(b), n, E.f64p(Tx), len(Tx),
                                         E.f64p(temp), n, 0, nb, 1, E.f64p(om), 1, blocksize))


# --------------------------------------------------------------------------------------------
# the smoothers sharing the SpMV / row-sweep core.  All of them go through the generic
# `amgb_host_relax` entry point (one smoother descriptor applied to host vectors by the cycle's own kernels).
# --------------------------------------------------------------------------------------------
def _relax(A, x, b, S, keep):
    """Apply the engine smoother descriptor S once: HOST A, x (in place), b."""
    E.require_gpu()
    M = E.as_matrix(A, keep)
    b = np.ascontiguousarray(b)
    E.check(E.lib().amgb_host_relax(M, S, E.f64p(x), E.f64p(b)))


def _descriptor():
    S = E.Smoother()
    S.kind, S.iterations, S.sweep, S.blocksize, S.omega = E.SM_NONE, 1, 0, 1, 1.0
    S.indices, S.n_indices, S.Dinv = None, 0, None
    S.indices2, S.n_indices2, S.f_iterations, S.c_iterations = None, 0, 1, 1
    S.coefficients, S.n_coefficients, S.reserved_ = None, 0, 0
    return S


def polynomial(A, x, b, coefficients, iterations=1):
    """x += p(A) (b - A x) with p given by its coefficients in descending order, evaluated by Horner's rule
    (relaxation.py:585-659); what the 'richardson' and 'chebyshev' smoothers call."""
    A, x, b = make_system(A, x, b, formats=None)
    _fp64(A)
    if A.shape[0] == 0 or iterations < 1:
        return
    coef = np.ascontiguousarray(np.real(np.asarray(coefficients)), dtype=np.float64).reshape(-1)
    if coef.size < 1:
        raise ValueError("polynomial smoother without coefficients")
    S = _descriptor()
    S.kind, S.iterations = E.SM_POLYNOMIAL, int(iterations)
    S.coefficients, S.n_coefficients = E.f64p(coef), coef.size
    _relax(A, x, b, S, [coef])


def jacobi_indexed(A, x, b, indices, iterations=1, omega=1.0):
    """Weighted Jacobi on the listed rows only (relaxation.py:1081-1138 -> relaxation.h:382-427): every listed
    row is relaxed from the iterate as it was when the call started."""
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _fp64(A)
    indices = np.ascontiguousarray(np.asarray(indices, dtype="intc"), dtype=np.int32)
    if A.format != "csr":
        R, C = A.blocksize
        if R != C:
            raise ValueError("BSR blocks must be square")
        if (R, C) != (1, 1):
            raise NotImplementedError("bsr_jacobi_indexed (block rows) is not on the GPU hot path")
        A = A.tocsr()
    n = A.shape[0]
    if n == 0 or len(indices) == 0:
        return
    E.require_gpu()
    L = E.lib()
    om = np.array([omega], dtype=np.float64)
    b = np.ascontiguousarray(b)
    Ap, Aj, Ax = _csr_args(A)
    for _ in range(iterations):
        E.check(L.amgb_host_jacobi_indexed(E.i32p(Ap), len(Ap), E.i32p(Aj), len(Aj), E.f64p(Ax), len(Ax),
                                           E.f64p(x), n, E.f64p(b), n, E.i32p(indices), len(indices),
                                           E.f64p(om), 1))


def _cf(kind, A, x, b, Cpts, Fpts, iterations, f_iterations, c_iterations, omega):
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _fp64(A)
    if A.format != "csr":
        R, C = A.blocksize
        if R != C:
            raise ValueError("BSR blocks must be square")
        if (R, C) != (1, 1):
            raise NotImplementedError("CF Jacobi on block rows (bsr_jacobi_indexed) is not on the GPU hot path")
        A = A.tocsr()
    Cpts = np.ascontiguousarray(np.asarray(Cpts), dtype=np.int32)
    Fpts = np.ascontiguousarray(np.asarray(Fpts), dtype=np.int32)
    if A.shape[0] == 0 or iterations < 1:
        return
    S = _descriptor()
    S.kind, S.iterations, S.omega = kind, int(iterations), float(np.real(omega))
    S.indices, S.n_indices = E.i32p(Cpts), len(Cpts)
    S.indices2, S.n_indices2 = E.i32p(Fpts), len(Fpts)
    S.f_iterations, S.c_iterations = int(f_iterations), int(c_iterations)
    _relax(A, x, b, S, [Cpts, Fpts])


def cf_jacobi(A, x, b, Cpts, Fpts, iterations=1, f_iterations=1, c_iterations=1, omega=1.0):
    """CF Jacobi: per iteration c_iterations Jacobi sweeps over the C-points, then f_iterations over the
    F-points (relaxation.py:1141-1203)."""
    _cf(E.SM_CF_JACOBI, A, x, b, Cpts, Fpts, iterations, f_iterations, c_iterations, omega)


def fc_jacobi(A, x, b, Cpts, Fpts, iterations=1, f_iterations=1, c_iterations=1, omega=1.0):
    """FC Jacobi: F-point sweeps first, then C-point sweeps (relaxation.py:1206-1268)."""
    _cf(E.SM_FC_JACOBI, A, x, b, Cpts, Fpts, iterations, f_iterations, c_iterations, omega)


def block_gauss_seidel(A, x, b, iterations=1, sweep="forward", blocksize=1, Dinv=None):
    """Block Gauss-Seidel, in place on x (relaxation.py:502-582 -> relaxation.h:1242-1298).  The sequential sweep
    over block rows runs as dependency waves of the block graph, which reproduces it."""
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _fp64(A)
    A = A.tobsr(blocksize=(blocksize, blocksize))
    if Dinv is None:
        from ..util import get_block_diag
        Dinv = get_block_diag(A, blocksize=blocksize, inv_flag=True)
    elif Dinv.shape[0] != int(A.shape[0] / blocksize):
        raise ValueError("Dinv and A have incompatible dimensions")
    elif (Dinv.shape[1] != blocksize) or (Dinv.shape[2] != blocksize):
        raise ValueError("Dinv and blocksize are incompatible")
    nb = int(len(x) / blocksize)
    if sweep == "forward":
        rs = (0, nb, 1)
    elif sweep == "backward":
        rs = (nb - 1, -1, -1)
    elif sweep == "symmetric":
        for _ in range(iterations):
            block_gauss_seidel(A, x, b, iterations=1, sweep="forward", blocksize=blocksize, Dinv=Dinv)
            block_gauss_seidel(A, x, b, iterations=1, sweep="backward", blocksize=blocksize, Dinv=Dinv)
        return
    else:
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
    if nb <= 0:
        return
    E.require_gpu()
    L = E.lib()
    n = A.shape[0]
    b = np.ascontiguousarray(b)
    Ap, Aj, Ax = _csr_args(A)
    Tx = np.ascontiguousarray(Dinv, dtype=np.float64).reshape(-1)
    for _ in range(iterations):
        E.check(L.amgb_host_block_gauss_seidel(E.i32p(Ap), len(Ap), E.i32p(Aj), len(Aj), E.f64p(Ax), len(Ax),
                                               E.f64p(x), n, E.f64p(b), n, E.f64p(Tx), len(Tx), *rs, blocksize))


def _cf_block(kind, A, x, b, Cpts, Fpts, Dinv, blocksize, iterations, f_iterations, c_iterations, omega):
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _fp64(A)
    A = A.tobsr(blocksize=(blocksize, blocksize))
    if Dinv is None:
        from ..util import get_block_diag
        Dinv = get_block_diag(A, blocksize=blocksize, inv_flag=True)
    elif Dinv.shape[0] != int(A.shape[0] / blocksize):
        raise ValueError("Dinv and A have incompatible dimensions")
    elif (Dinv.shape[1] != blocksize) or (Dinv.shape[2] != blocksize):
        raise ValueError("Dinv and blocksize are incompatible")
    Cpts = np.ascontiguousarray(np.asarray(Cpts), dtype=np.int32)
    Fpts = np.ascontiguousarray(np.asarray(Fpts), dtype=np.int32)
    if A.shape[0] == 0 or iterations < 1:
        return
    Tx = np.ascontiguousarray(Dinv, dtype=np.float64).reshape(-1)
    S = _descriptor()
    S.kind, S.iterations, S.omega, S.blocksize = kind, int(iterations), float(np.real(omega)), int(blocksize)
    S.Dinv = E.f64p(Tx)
    S.indices, S.n_indices = E.i32p(Cpts), len(Cpts)
    S.indices2, S.n_indices2 = E.i32p(Fpts), len(Fpts)
    S.f_iterations, S.c_iterations = int(f_iterations), int(c_iterations)
    _relax(A, x, b, S, [Cpts, Fpts, Tx])


def cf_block_jacobi(A, x, b, Cpts, Fpts, Dinv=None, blocksize=1, iterations=1, f_iterations=1, c_iterations=1,
                    omega=1.0):
    """CF block Jacobi: block rows listed in Cpts, then those in Fpts, each relaxed from a snapshot of the iterate
    (relaxation.py:1271-1339 -> block_jacobi_indexed, relaxation.h:1113-1172)."""
    _cf_block(E.SM_CF_BLOCK_JACOBI, A, x, b, Cpts, Fpts, Dinv, blocksize, iterations, f_iterations, c_iterations, omega)


def fc_block_jacobi(A, x, b, Cpts, Fpts, Dinv=None, blocksize=1, iterations=1, f_iterations=1, c_iterations=1,
                    omega=1.0):
    """FC block Jacobi (relaxation.py:1342-1412): F block rows first, then C."""
    _cf_block(E.SM_FC_BLOCK_JACOBI, A, x, b, Cpts, Fpts, Dinv, blocksize, iterations, f_iterations, c_iterations, omega)


def _normal_equations(kind, norm_eq, A, x, b, iterations, sweep, omega, Dinv):
    if not isinstance(x, np.ndarray):
        raise ValueError("expected numpy array for argument x")
    A, x, b = make_system(sparse.csr_array(A) if not (sparse.issparse(A) and A.format in ("csr", "bsr")) else A, x, b,
                          formats=["csr"])
    _fp64(A)
    if sweep not in E.SWEEPS:
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
    if A.shape[0] == 0 or iterations < 1:
        return
    if Dinv is None:
        from ..util import get_diagonal
        Dinv = get_diagonal(A, norm_eq=norm_eq, inv=True)
    Dinv = np.ascontiguousarray(np.ravel(Dinv), dtype=np.float64)
    S = _descriptor()
    S.kind, S.iterations, S.omega, S.sweep = kind, int(iterations), float(np.real(omega)), E.SWEEPS[sweep]
    S.Dinv = E.f64p(Dinv)
    _relax(A, x, b, S, [Dinv])


def jacobi_ne(A, x, b, iterations=1, omega=1.0):
    """Jacobi on A A^H y = b, x = A^H y (relaxation.py:734-812 -> relaxation.h:579-606):
    x += omega A^H (diag(A A^H)^-1 (b - A x))."""
    _normal_equations(E.SM_JACOBI_NE, 2, A, x, b, iterations, "forward", omega, None)


def gauss_seidel_ne(A, x, b, iterations=1, sweep="forward", omega=1.0, Dinv=None):
    """Gauss-Seidel on A A^H y = b (Kaczmarz row projections; relaxation.py:815-901 -> relaxation.h:633-657).  The
    sequential sweep runs as dependency waves of the rows' column-conflict graph, which reproduces it."""
    _normal_equations(E.SM_GAUSS_SEIDEL_NE, 2, A, x, b, iterations, sweep, omega, Dinv)


def gauss_seidel_nr(A, x, b, iterations=1, sweep="forward", omega=1.0, Dinv=None):
    """Gauss-Seidel on A^H A x = A^H b (column projections on the residual; relaxation.py:904-999 ->
    relaxation.h:684-713)."""
    _normal_equations(E.SM_GAUSS_SEIDEL_NR, 1, A, x, b, iterations, sweep, omega, Dinv)


# --------------------------------------------------------------------------------------------
# overlapping multiplicative Schwarz
# --------------------------------------------------------------------------------------------
def schwarz_parameters(A, subdomain=None, subdomain_ptr=None, inv_subblock=None, inv_subblock_ptr=None):
    """Subdomains and the (pseudo-)inverses of their diagonal blocks (relaxation.py:1002-1078).  HOST setup, cached
    on the matrix like the reference does: by default subdomain i is the sparsity pattern of row i; block i is
    A[sub_i][:, sub_i] (what amg_core.extract_subblocks gathers, relaxation.h:905-960) and its inverse the LAPACK
    gelss minimum-norm solve against the identity with rcond = 1e6 eps."""
    cached = getattr(A, "schwarz_parameters", None)
    if cached is not None:
        # the reference returns the cached set when no subdomains are named or the named ones are the cached ones
        if subdomain is None or subdomain_ptr is None or \
                (np.array_equal(cached[0], subdomain) and np.array_equal(cached[1], subdomain_ptr)):
            return cached
    if subdomain is None or subdomain_ptr is None:
        subdomain_ptr, subdomain = A.indptr.copy(), A.indices.copy()
    if inv_subblock is None or inv_subblock_ptr is None:
        from scipy.linalg import get_lapack_funcs
        sizes = np.diff(subdomain_ptr)
        inv_subblock_ptr = np.zeros(subdomain_ptr.shape, dtype=A.indices.dtype)
        np.cumsum(sizes * sizes, out=inv_subblock_ptr[1:])
        inv_subblock = np.zeros(int(inv_subblock_ptr[-1]), dtype=A.dtype)
        gelss, = get_lapack_funcs(["gelss"], (np.ones(1, dtype=A.dtype),))
        rcond = 1e6 * np.finfo(np.float64).eps
        Ap, Aj, Ax = A.indptr, A.indices, A.data
        where = np.full(A.shape[1], -1, dtype=np.int64)         # column -> local index inside the current subdomain
        for d, m in enumerate(sizes):
            rows = subdomain[subdomain_ptr[d]:subdomain_ptr[d + 1]]
            where[rows] = np.arange(m)
            block = np.zeros((m, m), dtype=A.dtype)
            for local, row in enumerate(rows):
                cols = Aj[Ap[row]:Ap[row + 1]]
                hit = where[cols] >= 0
                np.add.at(block[local], where[cols[hit]], Ax[Ap[row]:Ap[row + 1]][hit])
            where[rows] = -1
            sol = gelss(block, np.eye(m, dtype=A.dtype), cond=rcond, overwrite_a=True, overwrite_b=True)[1]
            inv_subblock[inv_subblock_ptr[d]:inv_subblock_ptr[d + 1]] = sol.reshape(-1)
    A.schwarz_parameters = (subdomain, subdomain_ptr, inv_subblock, inv_subblock_ptr)
    return A.schwarz_parameters


def schwarz(A, x, b, iterations=1, subdomain=None, subdomain_ptr=None, inv_subblock=None, inv_subblock_ptr=None,
            sweep="forward"):
    """Overlapping multiplicative Schwarz (relaxation.py:157-262 -> overlapping_schwarz_csr, relaxation.h:818-880):
    subdomain after subdomain, x[sub] += inv(A[sub, sub]) (b - A x)[sub].  On the device the subdomains run in
    dependency waves that reproduce the sequential order bit for bit."""
    A, x, b = make_system(A, x, b, formats=["csr"])
    _fp64(A)
    A.sort_indices()
    if sweep not in E.SWEEPS:
        raise ValueError("valid sweep directions: 'forward', 'backward', and 'symmetric'")
    subdomain, subdomain_ptr, inv_subblock, inv_subblock_ptr = schwarz_parameters(
        A, subdomain, subdomain_ptr, inv_subblock, inv_subblock_ptr)
    if A.shape[0] == 0 or iterations < 1:
        return
    S = _descriptor()
    keep = []
    _fill_schwarz(S, keep, subdomain, subdomain_ptr, inv_subblock, iterations, sweep)
    _relax(A, x, b, S, keep)


def _fill_schwarz(S, keep, subdomain, subdomain_ptr, inv_subblock, iterations, sweep):
    Sj = np.ascontiguousarray(subdomain, dtype=np.int32)
    Sp = np.ascontiguousarray(subdomain_ptr, dtype=np.int32)
    T = np.ascontiguousarray(np.real(inv_subblock), dtype=np.float64)
    keep += [Sj, Sp, T]
    S.kind, S.iterations, S.sweep = E.SM_SCHWARZ, int(iterations), E.SWEEPS[sweep]
    S.indices, S.n_indices = E.i32p(Sj), len(Sj)
    S.indices2, S.n_indices2 = E.i32p(Sp), len(Sp)
    S.Dinv = E.f64p(T)
