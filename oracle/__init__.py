"""CPU oracle for the AMG solve-phase hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product (pyamg_b200/) never does and has no CPU fallback.

Two layers, both restating the reference (paths relative to /root/reference/):

* native sweeps + matvecs: oracle/amg_oracle.c through ctypes (``kernels='oracle'``), or the
  reference's own relaxation.h compiled as oracle/_ref/libamg_ref.so (``kernels='ref'``;
  its SpMV leg is SciPy's csr_matvec/bsr_matvec -- exactly what the reference calls);
* Python wrappers with the signatures of pyamg/relaxation/relaxation.py (``jacobi`` :349-420,
  ``gauss_seidel`` :265-346, ``gauss_seidel_indexed`` :662-731, ``block_jacobi`` :423-499,
  ``make_system`` :15-97) and the cycle driver of pyamg/multilevel.py (``solve`` :398-582,
  ``__solve`` :584-662, 'pinv' coarse solver :717-721).

Parity status: PINNED -- see tests/test_oracle.py (reference KATs, golden fixtures produced
by the real reference via tests/golden/make_golden.py, and libamg_ref.so cross-checks).
"""
import ctypes
import os

import numpy as np
import scipy.sparse as sparse
from scipy.linalg import pinv

from .build import build_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_I = ctypes.POINTER(ctypes.c_int)
_D = ctypes.POINTER(ctypes.c_double)
_libs = {}


def _ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_I)


def _dp(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_D)


def lib(kind="oracle"):
    """Load (building if needed) the oracle C library, or the compiled reference shim."""
    if kind not in _libs:
        paths = build_oracle()
        p = paths["oracle" if kind == "oracle" else "ref"]
        if p is None:
            raise FileNotFoundError("oracle/_ref/libamg_ref.so not built (needs /root/reference)")
        _libs[kind] = ctypes.CDLL(p)
    return _libs[kind]


def have_ref():
    return build_oracle()["ref"] is not None


# --------------------------------------------------------------------------------------
# matvecs
# --------------------------------------------------------------------------------------
def matvec(A, x, kernels="oracle"):
    """A @ x the way the reference gets it (SciPy sparsetools), or the C restatement."""
    if kernels == "ref":
        return A @ x
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(A.shape[0], dtype=np.float64)
    L = lib("oracle")
    if A.format == "csr":
        L.oracle_csr_matvec(ctypes.c_int(A.shape[0]), _ip(A.indptr), _ip(A.indices),
                            _dp(A.data), _dp(x), _dp(y))
    elif A.format == "bsr":
        R, C = A.blocksize
        data = np.ascontiguousarray(A.data).ravel()
        L.oracle_bsr_matvec(ctypes.c_int(A.shape[0] // R), ctypes.c_int(R), ctypes.c_int(C),
                            _ip(A.indptr), _ip(A.indices), _dp(data), _dp(x), _dp(y))
    else:
        raise ValueError("matvec: csr or bsr expected")
    return y


# --------------------------------------------------------------------------------------
# relaxation wrappers (signatures of pyamg/relaxation/relaxation.py)
# --------------------------------------------------------------------------------------
def make_system(A, x, b, formats=None):
    """pyamg/relaxation/relaxation.py:15-97 (validation contract)."""
    if formats is None:
        pass
    elif formats == ["csr"]:
        if sparse.issparse(A) and A.format == "csr":
            pass
        elif sparse.issparse(A) and A.format == "bsr":
            A = A.tocsr()
        else:
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


def _f64(*arrs):
    for a in arrs:
        if a.dtype != np.float64:
            raise TypeError("oracle is fp64-only (BASELINE.json: fp64)")


def jacobi(A, x, b, iterations=1, omega=1.0, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:349-420."""
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _f64(A, x, b)
    n = A.shape[0]
    if n == 0:
        return
    temp = np.empty_like(x)
    om = ctypes.c_double(float(omega))
    if A.format == "csr":
        for _ in range(iterations):
            if kernels == "ref":
                lib("ref").ref_jacobi(_ip(A.indptr), n, _ip(A.indices), _dp(A.data), A.nnz,
                                      _dp(x), _dp(b), _dp(temp), 0, n, 1, om)
            else:
                lib().oracle_jacobi(_ip(A.indptr), _ip(A.indices), _dp(A.data), _dp(x), _dp(b),
                                    _dp(temp), 0, n, 1, om)
    else:
        R, C = A.blocksize
        if R != C:
            raise ValueError("BSR blocks must be square")
        data = np.ascontiguousarray(A.data).ravel()
        nb = n // R
        for _ in range(iterations):
            if kernels == "ref":
                lib("ref").ref_bsr_jacobi(_ip(A.indptr), nb, _ip(A.indices), _dp(data),
                                          len(A.indices), _dp(x), _dp(b), _dp(temp),
                                          0, nb, 1, R, om)
            else:
                lib().oracle_bsr_jacobi(_ip(A.indptr), _ip(A.indices), _dp(data), _dp(x),
                                        _dp(b), _dp(temp), 0, nb, 1, R, om)


def gauss_seidel(A, x, b, iterations=1, sweep="forward", omega=1.0, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:265-346 (CSR path; 'symmetric' = fwd then bwd WITHOUT omega,
    :326-330; omega != 1 -> sor_gauss_seidel :335-338)."""
    if sparse.issparse(A) and A.format == "bsr":
        # relaxation.py:343-346: the BSR branch calls bsr_gauss_seidel, which has no omega -- SOR on a BSR operator
        # (every smoothed-aggregation coarse level is BSR, blocksize 1 for scalar problems) is plain Gauss-Seidel;
        # point-wise BSR Gauss-Seidel equals the sweep on the CSR expansion (test_relaxation.py:224-249)
        omega = 1.0
    A, x, b = make_system(A, x, b, formats=["csr"])
    _f64(A, x, b)
    n = A.shape[0]
    if sweep == "forward":
        rs = (0, n, 1)
    elif sweep == "backward":
        rs = (n - 1, -1, -1)
    elif sweep == "symmetric":
        for _ in range(iterations):
            gauss_seidel(A, x, b, iterations=1, sweep="forward", kernels=kernels)
            gauss_seidel(A, x, b, iterations=1, sweep="backward", kernels=kernels)
        return
    else:
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
    for _ in range(iterations):
        if omega != 1.0:
            lib().oracle_sor_gauss_seidel(_ip(A.indptr), _ip(A.indices), _dp(A.data), _dp(x),
                                          _dp(b), *rs, ctypes.c_double(float(omega)))
        elif kernels == "ref":
            lib("ref").ref_gauss_seidel(_ip(A.indptr), n, _ip(A.indices), _dp(A.data), A.nnz,
                                        _dp(x), _dp(b), *rs)
        else:
            lib().oracle_gauss_seidel(_ip(A.indptr), _ip(A.indices), _dp(A.data), _dp(x),
                                      _dp(b), *rs)


def sor(A, x, b, omega, iterations=1, sweep="forward", kernels="oracle"):
    """pyamg/relaxation/relaxation.py:100-154."""
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    for _ in range(iterations):
        gauss_seidel(A, x, b, iterations=1, sweep=sweep, omega=omega, kernels=kernels)


def gauss_seidel_indexed(A, x, b, indices, iterations=1, sweep="forward", kernels="oracle"):
    """pyamg/relaxation/relaxation.py:662-731."""
    A, x, b = make_system(A, x, b, formats=["csr"])
    _f64(A, x, b)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    m = len(indices)
    if sweep == "forward":
        rs = (0, m, 1)
    elif sweep == "backward":
        rs = (m - 1, -1, -1)
    elif sweep == "symmetric":
        for _ in range(iterations):
            gauss_seidel_indexed(A, x, b, indices, 1, "forward", kernels)
            gauss_seidel_indexed(A, x, b, indices, 1, "backward", kernels)
        return
    else:
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
    n = A.shape[0]
    for _ in range(iterations):
        if kernels == "ref":
            lib("ref").ref_gauss_seidel_indexed(_ip(A.indptr), n, _ip(A.indices), _dp(A.data),
                                                A.nnz, _dp(x), _dp(b), _ip(indices), m, *rs)
        else:
            lib().oracle_gauss_seidel_indexed(_ip(A.indptr), _ip(A.indices), _dp(A.data),
                                              _dp(x), _dp(b), _ip(indices), *rs)


def jacobi_indexed(A, x, b, indices, iterations=1, omega=1.0, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:1081-1138 -> relaxation.h:382-427."""
    A, x, b = make_system(A, x, b, formats=["csr"])
    _f64(A, x, b)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    for _ in range(iterations):
        if kernels == "ref":
            lib("ref").ref_jacobi_indexed(_ip(A.indptr), A.shape[0], _ip(A.indices), _dp(A.data), len(A.data),
                                          _dp(x), _dp(b), _ip(indices), len(indices), ctypes.c_double(float(omega)))
        else:
            lib().oracle_jacobi_indexed(_ip(A.indptr), _ip(A.indices), _dp(A.data), _dp(x),
                                        A.shape[0], _dp(b), _ip(indices), len(indices),
                                        ctypes.c_double(float(omega)))


def cf_jacobi(A, x, b, Cpts, Fpts, iterations=1, f_iterations=1, c_iterations=1, omega=1.0, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:1141-1203 (CSR branch): C sweeps, then F sweeps."""
    for _ in range(iterations):
        for _c in range(c_iterations):
            jacobi_indexed(A, x, b, Cpts, omega=omega, kernels=kernels)
        for _f in range(f_iterations):
            jacobi_indexed(A, x, b, Fpts, omega=omega, kernels=kernels)


def fc_jacobi(A, x, b, Cpts, Fpts, iterations=1, f_iterations=1, c_iterations=1, omega=1.0, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:1206-1268 (CSR branch): F sweeps, then C sweeps."""
    for _ in range(iterations):
        for _f in range(f_iterations):
            jacobi_indexed(A, x, b, Fpts, omega=omega, kernels=kernels)
        for _c in range(c_iterations):
            jacobi_indexed(A, x, b, Cpts, omega=omega, kernels=kernels)


def _block_jacobi_indexed(A, x, b, Dinv, indices, blocksize, omega, kernels):
    nb = A.shape[0] // blocksize
    data = np.ascontiguousarray(A.data).ravel()
    dinv = np.ascontiguousarray(Dinv, dtype=np.float64).ravel()
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    om = ctypes.c_double(float(omega))
    if kernels == "ref":
        lib("ref").ref_block_jacobi_indexed(_ip(A.indptr), nb, _ip(A.indices), _dp(data), len(A.indices), _dp(x), _dp(b),
                                            _dp(dinv), _ip(indices), len(indices), om, blocksize)
    else:
        lib().oracle_block_jacobi_indexed(_ip(A.indptr), _ip(A.indices), _dp(data), _dp(x), A.shape[0], _dp(b),
                                          _dp(dinv), _ip(indices), len(indices), om, blocksize)


def _cf_block_jacobi(order, A, x, b, Cpts, Fpts, Dinv, blocksize, iterations, f_iterations, c_iterations, omega, kernels):
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _f64(A, x, b)
    A = A.tobsr(blocksize=(blocksize, blocksize))
    if Dinv is None:
        raise ValueError("oracle: block CF Jacobi needs Dinv (setup-time quantity)")
    for _ in range(iterations):
        for pts, reps in ((Cpts, c_iterations), (Fpts, f_iterations)) if order == "cf" else ((Fpts, f_iterations), (Cpts, c_iterations)):
            for _r in range(reps):
                _block_jacobi_indexed(A, x, b, Dinv, pts, blocksize, omega, kernels)


def cf_block_jacobi(A, x, b, Cpts, Fpts, Dinv=None, blocksize=1, iterations=1, f_iterations=1, c_iterations=1,
                    omega=1.0, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:1271-1339."""
    _cf_block_jacobi("cf", A, x, b, Cpts, Fpts, Dinv, blocksize, iterations, f_iterations, c_iterations, omega, kernels)


def fc_block_jacobi(A, x, b, Cpts, Fpts, Dinv=None, blocksize=1, iterations=1, f_iterations=1, c_iterations=1,
                    omega=1.0, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:1342-1412."""
    _cf_block_jacobi("fc", A, x, b, Cpts, Fpts, Dinv, blocksize, iterations, f_iterations, c_iterations, omega, kernels)


def get_diagonal(A, norm_eq=False, inv=False):
    """pyamg/util/utils.py:530-600: diag(A), diag(A^H A) (norm_eq=1) or diag(A A^H) (norm_eq=2); sorts A IN PLACE."""
    A.sort_indices()
    if norm_eq == 1:
        At = A.T
        D = (At.multiply(At.conjugate())) @ np.ones((At.shape[0],))
    elif norm_eq == 2:
        D = (A.multiply(A.conjugate())) @ np.ones((A.shape[0],))
    else:
        D = A.diagonal()
    if inv:
        Dinv = np.zeros_like(D)
        mask = D != 0.0
        Dinv[mask] = 1.0 / D[mask]
        return Dinv
    return D


def jacobi_ne(A, x, b, iterations=1, omega=1.0, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:734-812 -> relaxation.h:579-606."""
    A, x, b = make_system(A, x, b, formats=["csr"])
    _f64(A, x, b)
    n = A.shape[0]
    temp = np.zeros_like(x)
    Dinv = get_diagonal(A, norm_eq=2, inv=True)
    for _ in range(iterations):
        delta = np.ascontiguousarray(np.ravel(b - matvec(A, x, kernels)) * np.ravel(Dinv))
        if kernels == "ref":
            lib("ref").ref_jacobi_ne(_ip(A.indptr), n, _ip(A.indices), _dp(A.data), len(A.data), _dp(x), _dp(b),
                                     _dp(delta), _dp(temp), ctypes.c_double(float(omega)))
        else:
            lib().oracle_jacobi_ne(_ip(A.indptr), _ip(A.indices), _dp(A.data), _dp(x), _dp(delta), _dp(temp), n,
                                   ctypes.c_double(float(omega)))


def gauss_seidel_ne(A, x, b, iterations=1, sweep="forward", omega=1.0, Dinv=None, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:815-901 -> relaxation.h:633-657."""
    A, x, b = make_system(A, x, b, formats=["csr"])
    _f64(A, x, b)
    n = A.shape[0]
    if Dinv is None:
        Dinv = np.ravel(get_diagonal(A, norm_eq=2, inv=True))
    if sweep == "forward":
        rs = (0, n, 1)
    elif sweep == "backward":
        rs = (n - 1, -1, -1)
    elif sweep == "symmetric":
        for _ in range(iterations):
            gauss_seidel_ne(A, x, b, 1, "forward", omega, Dinv, kernels)
            gauss_seidel_ne(A, x, b, 1, "backward", omega, Dinv, kernels)
        return
    else:
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
    Dinv = np.ascontiguousarray(Dinv, dtype=np.float64)
    for _ in range(iterations):
        if kernels == "ref":
            lib("ref").ref_gauss_seidel_ne(_ip(A.indptr), n, _ip(A.indices), _dp(A.data), len(A.data), _dp(x), _dp(b),
                                           *rs, _dp(Dinv), ctypes.c_double(float(omega)))
        else:
            lib().oracle_gauss_seidel_ne(_ip(A.indptr), _ip(A.indices), _dp(A.data), _dp(x), _dp(b), *rs, _dp(Dinv),
                                         ctypes.c_double(float(omega)))


def gauss_seidel_nr(A, x, b, iterations=1, sweep="forward", omega=1.0, Dinv=None, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:904-999 -> relaxation.h:684-713 (A in CSC)."""
    if not (sparse.issparse(A) and A.format == "csc"):
        A = sparse.csc_array(A)
    A.indptr, A.indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    if not isinstance(x, np.ndarray) or not isinstance(b, np.ndarray):
        raise ValueError("expected numpy arrays")
    x, b = np.ravel(x), np.ravel(b)
    n = A.shape[0]
    if Dinv is None:
        Dinv = np.ravel(get_diagonal(A, norm_eq=1, inv=True))
    if sweep == "forward":
        cs = (0, n, 1)
    elif sweep == "backward":
        cs = (n - 1, -1, -1)
    elif sweep == "symmetric":
        for _ in range(iterations):
            gauss_seidel_nr(A, x, b, 1, "forward", omega, Dinv, kernels)
            gauss_seidel_nr(A, x, b, 1, "backward", omega, Dinv, kernels)
        return
    else:
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
    Dinv = np.ascontiguousarray(Dinv, dtype=np.float64)
    r = np.ascontiguousarray(b - A @ x)
    data = np.ascontiguousarray(A.data, dtype=np.float64)
    for _ in range(iterations):
        if kernels == "ref":
            lib("ref").ref_gauss_seidel_nr(_ip(A.indptr), n, _ip(A.indices), _dp(data), len(data), _dp(x), _dp(r),
                                           *cs, _dp(Dinv), ctypes.c_double(float(omega)))
        else:
            lib().oracle_gauss_seidel_nr(_ip(A.indptr), _ip(A.indices), _dp(data), _dp(x), _dp(r), *cs, _dp(Dinv),
                                         ctypes.c_double(float(omega)))


def schwarz_parameters(A, subdomain=None, subdomain_ptr=None, inv_subblock=None, inv_subblock_ptr=None):
    """pyamg/relaxation/relaxation.py:1002-1078: default subdomains = the rows' sparsity patterns; every subdomain
    block A[sub][:, sub] (amg_core.extract_subblocks, relaxation.h:905-960) is pseudo-inverted with LAPACK gelss."""
    from scipy.linalg import get_lapack_funcs
    if hasattr(A, "schwarz_parameters"):
        if subdomain is None or subdomain_ptr is None or (np.array_equal(A.schwarz_parameters[0], subdomain)
                                                          and np.array_equal(A.schwarz_parameters[1], subdomain_ptr)):
            return A.schwarz_parameters
    if subdomain is None or subdomain_ptr is None:
        subdomain_ptr = A.indptr.copy()
        subdomain = A.indices.copy()
    if inv_subblock is None or inv_subblock_ptr is None:
        inv_subblock_ptr = np.zeros(subdomain_ptr.shape, dtype=A.indices.dtype)
        blocksize = subdomain_ptr[1:] - subdomain_ptr[:-1]
        inv_subblock_ptr[1:] = np.cumsum(blocksize * blocksize)
        inv_subblock = np.zeros((inv_subblock_ptr[-1],), dtype=A.dtype)
        cond = 1e6 * np.finfo(np.double).eps                # util/params.py set_tol('d')
        gelss, = get_lapack_funcs(["gelss"], (np.ones((1,), dtype=A.dtype),))
        Acsr = sparse.csr_array(A)
        for i in range(subdomain_ptr.shape[0] - 1):
            m = blocksize[i]
            sub = subdomain[subdomain_ptr[i]:subdomain_ptr[i + 1]]
            block = np.array(Acsr[sub][:, sub].toarray(), order="C")
            out = gelss(block, np.eye(m, m, dtype=A.dtype), cond=cond, overwrite_a=True, overwrite_b=True)
            inv_subblock[inv_subblock_ptr[i]:inv_subblock_ptr[i + 1]] = np.ravel(out[1])
    A.schwarz_parameters = (subdomain, subdomain_ptr, inv_subblock, inv_subblock_ptr)
    return A.schwarz_parameters


def schwarz(A, x, b, iterations=1, subdomain=None, subdomain_ptr=None, inv_subblock=None, inv_subblock_ptr=None,
            sweep="forward", kernels="oracle"):
    """pyamg/relaxation/relaxation.py:157-262 -> overlapping_schwarz_csr (relaxation.h:818-880)."""
    A, x, b = make_system(A, x, b, formats=["csr"])
    _f64(A, x, b)
    A.sort_indices()
    subdomain, subdomain_ptr, inv_subblock, inv_subblock_ptr = schwarz_parameters(A, subdomain, subdomain_ptr,
                                                                                  inv_subblock, inv_subblock_ptr)
    nsub = subdomain_ptr.shape[0] - 1
    if sweep == "forward":
        rs = (0, nsub, 1)
    elif sweep == "backward":
        rs = (nsub - 1, -1, -1)
    elif sweep == "symmetric":
        for _ in range(iterations):
            schwarz(A, x, b, 1, subdomain, subdomain_ptr, inv_subblock, inv_subblock_ptr, "forward", kernels)
            schwarz(A, x, b, 1, subdomain, subdomain_ptr, inv_subblock, inv_subblock_ptr, "backward", kernels)
        return
    else:
        raise ValueError("valid sweep directions: 'forward', 'backward', and 'symmetric'")
    Sj = np.ascontiguousarray(subdomain, dtype=np.int32)
    Sp = np.ascontiguousarray(subdomain_ptr, dtype=np.int32)
    Tx = np.ascontiguousarray(inv_subblock, dtype=np.float64)
    Tp = np.ascontiguousarray(inv_subblock_ptr, dtype=np.int32)
    n = A.shape[0]
    for _ in range(iterations):
        if kernels == "ref":
            lib("ref").ref_overlapping_schwarz_csr(_ip(A.indptr), n, _ip(A.indices), _dp(A.data), len(A.data), _dp(x), _dp(b),
                                                   _dp(Tx), len(Tx), _ip(Tp), _ip(Sj), len(Sj), _ip(Sp), nsub, *rs)
        else:
            lib().oracle_overlapping_schwarz_csr(_ip(A.indptr), _ip(A.indices), _dp(A.data), _dp(x), _dp(b), _dp(Tx),
                                                 _ip(Tp), _ip(Sj), _ip(Sp), n, *rs)


def polynomial(A, x, b, coefficients, iterations=1, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:585-659: x += p(A)(b - A x) by Horner's rule; the matvecs are the
    reference's SciPy calls restated (``matvec``)."""
    A, x, b = make_system(A, x, b, formats=None)
    _f64(A, x, b)
    for _ in range(iterations):
        if np.linalg.norm(x) == 0:                       # :646-649
            residual = b
        else:
            residual = b - matvec(A, x, kernels)
        h = coefficients[0] * residual                   # :651
        for c in coefficients[1:]:                       # :653-654
            h = c * residual + matvec(A, h, kernels)
        x += h                                           # :656


def block_gauss_seidel(A, x, b, iterations=1, sweep="forward", blocksize=1, Dinv=None, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:502-582 -> relaxation.h:1242-1298 (Dinv must be given)."""
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _f64(A, x, b)
    A = A.tobsr(blocksize=(blocksize, blocksize))
    if Dinv is None:
        raise ValueError("oracle.block_gauss_seidel needs Dinv (setup-time quantity)")
    if Dinv.shape[0] != A.shape[0] // blocksize:
        raise ValueError("Dinv and A have incompatible dimensions")
    if Dinv.shape[1] != blocksize or Dinv.shape[2] != blocksize:
        raise ValueError("Dinv and blocksize are incompatible")
    nb = len(x) // blocksize
    if sweep == "forward":
        rs = (0, nb, 1)
    elif sweep == "backward":
        rs = (nb - 1, -1, -1)
    elif sweep == "symmetric":
        for _ in range(iterations):
            block_gauss_seidel(A, x, b, 1, "forward", blocksize, Dinv, kernels)
            block_gauss_seidel(A, x, b, 1, "backward", blocksize, Dinv, kernels)
        return
    else:
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
    if nb == 0:
        return
    data = np.ascontiguousarray(A.data).ravel()
    dinv = np.ascontiguousarray(Dinv, dtype=np.float64).ravel()
    for _ in range(iterations):
        if kernels == "ref":
            lib("ref").ref_block_gauss_seidel(_ip(A.indptr), nb, _ip(A.indices), _dp(data), len(A.indices),
                                              _dp(x), _dp(b), _dp(dinv), *rs, blocksize)
        else:
            lib().oracle_block_gauss_seidel(_ip(A.indptr), _ip(A.indices), _dp(data), _dp(x), _dp(b),
                                            _dp(dinv), *rs, blocksize)


def block_jacobi(A, x, b, Dinv=None, blocksize=1, iterations=1, omega=1.0, kernels="oracle"):
    """pyamg/relaxation/relaxation.py:423-499 (Dinv must be given: computing it is setup)."""
    A, x, b = make_system(A, x, b, formats=["csr", "bsr"])
    _f64(A, x, b)
    A = A.tobsr(blocksize=(blocksize, blocksize))
    if Dinv is None:
        raise ValueError("oracle.block_jacobi needs Dinv (setup-time quantity)")
    if Dinv.shape[0] != A.shape[0] // blocksize:
        raise ValueError("Dinv and A have incompatible dimensions")
    if Dinv.shape[1] != blocksize or Dinv.shape[2] != blocksize:
        raise ValueError("Dinv and blocksize are incompatible")
    nb = A.shape[0] // blocksize
    if nb == 0:
        return
    temp = np.empty_like(x)
    data = np.ascontiguousarray(A.data).ravel()
    dinv = np.ascontiguousarray(Dinv, dtype=np.float64).ravel()
    om = ctypes.c_double(float(omega))
    for _ in range(iterations):
        if kernels == "ref":
            lib("ref").ref_block_jacobi(_ip(A.indptr), nb, _ip(A.indices), _dp(data),
                                        len(A.indices), _dp(x), _dp(b), _dp(dinv), _dp(temp),
                                        0, nb, 1, om, blocksize)
        else:
            lib().oracle_block_jacobi(_ip(A.indptr), _ip(A.indices), _dp(data), _dp(x), _dp(b),
                                      _dp(dinv), _dp(temp), 0, nb, 1, om, blocksize)


_SMOOTHERS = {
    "jacobi": jacobi,
    "gauss_seidel": gauss_seidel,
    "gauss_seidel_indexed": gauss_seidel_indexed,
    "block_jacobi": block_jacobi,
    "sor": sor,
    "polynomial": polynomial,
    "jacobi_indexed": jacobi_indexed,
    "cf_jacobi": cf_jacobi,
    "fc_jacobi": fc_jacobi,
    "block_gauss_seidel": block_gauss_seidel,
    "cf_block_jacobi": cf_block_jacobi,
    "fc_block_jacobi": fc_block_jacobi,
    "jacobi_ne": jacobi_ne,
    "gauss_seidel_ne": gauss_seidel_ne,
    "gauss_seidel_nr": gauss_seidel_nr,
    "schwarz": schwarz,
}


# --------------------------------------------------------------------------------------
# hierarchy spec + cycle driver
# --------------------------------------------------------------------------------------
def smoother_spec(sm):
    """(function-name, kwargs) of a per-level smoother object, or None.

    Accepts what pyamg's change_smoothers stores (functools.partial with .func/.keywords,
    pyamg/relaxation/smoothing.py:494-579), the no-op `none` function (:833-837), or an
    already-neutral (name, kwargs) tuple.
    """
    if sm is None:
        return None
    if isinstance(sm, tuple):
        return (sm[0], dict(sm[1]))
    func = getattr(sm, "func", None)
    if func is not None:
        return (func.__name__, dict(sm.keywords))
    if getattr(sm, "__name__", "") == "none":
        return None
    if getattr(sm, "__name__", "") in ("richardson", "chebyshev") and getattr(sm, "__closure__", None):
        # smoothing.py:611-618 / :627-647: closures around relaxation.polynomial; parameters live in the cells
        cv = {k: c.cell_contents for k, c in zip(sm.__code__.co_freevars, sm.__closure__)}
        coef = cv["coefficients"] if "coefficients" in cv else [cv["omega"]]
        return ("polynomial", {"coefficients": np.asarray(coef, dtype=np.float64),
                               "iterations": int(cv.get("iterations", 1))})
    if getattr(sm, "__name__", "") == "schwarz" and getattr(sm, "__closure__", None):
        # smoothing.py:509-526: closure around relaxation.schwarz(lvl.Acsr, ...) with the subdomains and their block
        # inverses (computed once at setup) in the cells
        own = getattr(sm, "_schwarz_parameters", None)
        cv = own if own is not None else {k: c.cell_contents for k, c in zip(sm.__code__.co_freevars, sm.__closure__)}
        return ("schwarz", {k: cv[k] for k in ("iterations", "subdomain", "subdomain_ptr", "inv_subblock",
                                               "inv_subblock_ptr", "sweep")})
    if getattr(sm, "__name__", "") == "strength_based_schwarz" and getattr(sm, "__closure__", None):
        # smoothing.py:529-548: subdomains = rows of the strength matrix in the cells; the block inverses are rebuilt
        # from lvl.Acsr on every application (inv_subblock=None below does the same)
        own = getattr(sm, "_schwarz_parameters", None)      # closures built by pyamg_b200 carry them as an attribute
        if own is not None:
            return ("schwarz", {k: own[k] for k in ("iterations", "subdomain", "subdomain_ptr", "inv_subblock",
                                                    "inv_subblock_ptr", "sweep")})
        cv = {k: c.cell_contents for k, c in zip(sm.__code__.co_freevars, sm.__closure__)}
        return ("schwarz", {k: cv[k] for k in ("iterations", "subdomain", "subdomain_ptr", "sweep")})
    if getattr(sm, "__name__", "") in ("jacobi_ne", "gauss_seidel_ne", "gauss_seidel_nr") and getattr(sm, "__closure__", None):
        # smoothing.py:641-675: closures around relaxation.<name>(lvl.Acsr | lvl.Acsc, x, b, ...); the operator they
        # relax with is the level's own A in another format, so only the scalar parameters are taken from the cells
        own = getattr(sm, "_ne_parameters", None)          # closures built by pyamg_b200 carry them as an attribute
        if own is not None:
            return (sm.__name__, dict(own))
        cv = {k: c.cell_contents for k, c in zip(sm.__code__.co_freevars, sm.__closure__)}
        kw = {k: cv[k] for k in ("iterations", "sweep", "omega") if k in cv}
        return (sm.__name__, kw)
    raise NotImplementedError(f"oracle: smoother {sm!r} is not introspectable")


def hierarchy_spec(ml):
    """Neutral description of a MultilevelSolver-like object (pyamg's or pyamg_b200's)."""
    levels = []
    for lvl in ml.levels:
        d = {"A": lvl.A}
        if hasattr(lvl, "P"):
            d["P"] = lvl.P
            d["R"] = lvl.R if hasattr(lvl, "R") else lvl.P.T.conjugate()
            d["pre"] = smoother_spec(getattr(lvl, "presmoother", None))
            d["post"] = smoother_spec(getattr(lvl, "postsmoother", None))
        levels.append(d)
    # a relaxation method as the coarsest-level solver (multilevel.py:764-781): recorded on the last level
    cs = getattr(ml, "coarse_solver", None)
    if getattr(cs, "relaxation", None) is not None and levels:
        levels[-1]["coarse_relax"] = smoother_spec(cs.smoother(levels[-1]["A"]))
    return levels


def _smooth(spec, A, x, b, kernels):
    if spec is None:
        return
    name, kw = spec
    fn = _SMOOTHERS.get(name)
    if fn is None:
        raise NotImplementedError(f"oracle: smoother '{name}' out of hot-path scope")
    if name in ("jacobi_ne", "gauss_seidel_ne", "schwarz"):
        A = sparse.csr_array(A).copy()             # lvl.Acsr: a CSR view the reference sorts in place, not lvl.A
        A.indptr, A.indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    elif name == "gauss_seidel_nr":
        A = sparse.csc_array(sparse.csr_array(A))   # lvl.Acsc
    fn(A, x, b, kernels=kernels, **kw)


def cg(A, b, x0=None, tol=1e-5, maxiter=None, M=None, callback=None, residuals=None, kernels="oracle"):
    """pyamg/krylov/_cg.py:11-196 with criteria='rr' (the default the reference's solve(accel='cg') uses)."""
    b = np.ravel(np.asarray(b, dtype=np.float64))
    n = len(b)
    x = np.zeros(n) if x0 is None else np.array(np.ravel(x0), dtype=np.float64)
    if maxiter is None:
        maxiter = int(1.3 * n) + 2
    elif maxiter < 1:
        raise ValueError("Number of iterations must be positive")
    r = b - matvec(A, x, kernels)
    z = M(r)
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
        Ap = matvec(A, p, kernels)
        rz_old = rz
        pAp = np.inner(Ap, p)
        if pAp < 0.0:
            return x, -1
        alpha = rz / pAp
        x += alpha * p
        if np.mod(it, 8) and it > 0:
            r -= alpha * Ap
        else:
            r = b - matvec(A, x, kernels)
        z = M(r)
        rz = np.inner(r, z)
        if rz < 0.0:
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


class Cycle:
    """Restatement of MultilevelSolver.solve/__solve (pyamg/multilevel.py:398-662), V/W/F/AMLI cycles,
    'pinv' coarse solve (:717-721, GenericSolver.__call__ :797-816)."""

    def __init__(self, levels, coarse_pinv=None, kernels="oracle"):
        self.levels = levels
        self.kernels = kernels
        self.coarse_pinv = coarse_pinv

    def coarse_solve(self, A, b):
        if A.nnz == 0:
            return np.zeros(b.shape)
        relax = self.levels[-1].get("coarse_relax")
        if relax is not None:                      # multilevel.py:773-779: x = 0; relax(A, x, b)
            x = np.zeros_like(b)
            _smooth(relax, A, x, np.ascontiguousarray(b), self.kernels)
            return x
        if self.coarse_pinv is None:
            self.coarse_pinv = pinv(A.toarray())
        return np.dot(self.coarse_pinv, b)

    def solve(self, b, x0=None, tol=1e-5, maxiter=100, cycle="V", residuals=None,
              callback=None, cycles_per_level=1, return_info=False, accel=None):
        if accel == "cg":      # multilevel.py:479-508 with pyamg.krylov.cg; M = aspreconditioner (:390-396)
            x, info = cg(self.levels[0]["A"], b, x0=x0, tol=tol, maxiter=maxiter,
                         M=lambda r: self.solve(r, maxiter=1, cycle=cycle, tol=1e-12),
                         callback=callback, residuals=residuals, kernels=self.kernels)
            x = x.reshape(np.shape(b))
            return (x, info) if return_info else x
        if accel in ("gmres", "fgmres"):   # pyamg.krylov.gmres (Householder) / fgmres: oracle/krylov.py
            from .krylov import gmres as _gmres
            A0 = self.levels[0]["A"]
            x, info = _gmres(A0, b, x0=x0, tol=tol, maxiter=maxiter, flexible=(accel == "fgmres"),
                             M=lambda r: self.solve(r, maxiter=1, cycle=cycle, tol=1e-12),
                             matvec=lambda v: matvec(A0, v, self.kernels), callback=callback, residuals=residuals)
            x = x.reshape(np.shape(b))
            return (x, info) if return_info else x
        if accel == "bicgstab":            # pyamg.krylov.bicgstab: oracle/krylov.py
            from .krylov import bicgstab as _bicgstab
            A0 = self.levels[0]["A"]
            x, info = _bicgstab(A0, b, x0=x0, tol=tol, maxiter=maxiter,
                                M=lambda r: self.solve(r, maxiter=1, cycle=cycle, tol=1e-12),
                                matvec=lambda v: matvec(A0, v, self.kernels), callback=callback, residuals=residuals)
            x = x.reshape(np.shape(b))
            return (x, info) if return_info else x
        if accel is not None:
            raise NotImplementedError("oracle: accel other than 'cg', 'gmres', 'fgmres', 'bicgstab'")
        x = np.zeros_like(b) if x0 is None else np.array(x0)
        A = self.levels[0]["A"]
        cycle = str(cycle).upper()
        normb = np.linalg.norm(b)
        if normb == 0.0:
            normb = 1.0
        normr = np.linalg.norm(b - matvec(A, np.ravel(x), self.kernels).reshape(b.shape))
        if residuals is not None:
            residuals[:] = [normr]
        b = np.ravel(np.asarray(b, dtype=np.float64))
        x = np.ravel(np.asarray(x, dtype=np.float64))
        it = 0
        while True:
            if len(self.levels) == 1:
                x = self.coarse_solve(A, b)
            else:
                self._solve(0, x, b, cycle, cycles_per_level)
            it += 1
            normr = np.linalg.norm(b - matvec(A, x, self.kernels))
            if residuals is not None:
                residuals.append(normr)
            if callback is not None:
                callback(x)
            if normr < tol * normb:
                return (x, 0) if return_info else x
            if it == maxiter:
                return (x, it) if return_info else x

    def _solve(self, lvl, x, b, cycle, cycles_per_level=1):
        L = self.levels[lvl]
        A = L["A"]
        _smooth(L["pre"], A, x, b, self.kernels)
        residual = b - matvec(A, x, self.kernels)
        coarse_b = matvec(L["R"], residual, self.kernels)
        coarse_x = np.zeros_like(coarse_b)
        if lvl == len(self.levels) - 2:
            coarse_x[:] = self.coarse_solve(self.levels[-1]["A"], coarse_b)
        elif cycle == "V":
            self._solve(lvl + 1, coarse_x, coarse_b, "V")
        elif cycle == "W":
            self._solve(lvl + 1, coarse_x, coarse_b, cycle)
            self._solve(lvl + 1, coarse_x, coarse_b, cycle)
        elif cycle == "F":
            self._solve(lvl + 1, coarse_x, coarse_b, cycle, cycles_per_level)
            for _ in range(cycles_per_level):
                self._solve(lvl + 1, coarse_x, coarse_b, "V", 1)
        elif cycle == "AMLI":
            # multilevel.py:631-657: two A_c-orthogonalised coarse corrections; every inner solve starts from an
            # all-ones guess and sees the coarse rhs as updated by the previous step
            nAMLI = 2
            Ac = self.levels[lvl + 1]["A"]
            p = np.zeros((nAMLI, coarse_b.shape[0]))
            for k in range(nAMLI):
                p[k, :] = 1
                self._solve(lvl + 1, p[k, :], coarse_b, cycle)
                for j in range(k):
                    beta = np.inner(p[j, :], matvec(Ac, p[k, :], self.kernels)) / \
                        np.inner(p[j, :], matvec(Ac, p[j, :], self.kernels))
                    p[k, :] -= beta * p[j, :]
                Ap = matvec(Ac, p[k, :], self.kernels)
                alpha = np.inner(p[k, :], coarse_b) / np.inner(p[k, :], Ap)
                coarse_x += alpha * p[k, :]
                coarse_b -= alpha * Ap
        else:
            raise TypeError(f"Unrecognized cycle type ({cycle})")
        x += matvec(L["P"], coarse_x, self.kernels)
        _smooth(L["post"], A, x, b, self.kernels)
