"""Setup-time helpers the smoother factory needs (host, NumPy/SciPy; run once per hierarchy).

These are NOT on the hot path.  When a hierarchy comes from the reference, omega and Dinv are
read from its smoother closures (SURVEY.md hazard 2) and nothing here runs.  They exist so that
``pyamg_b200.relaxation.smoothing.change_smoothers`` can equip hierarchies built without the
reference (bench inputs on the GPU box) with the same quantities:
  get_diagonal                 <-> pyamg/util/utils.py:541-600
  get_block_diag               <-> pyamg/util/utils.py:603-692
  approximate_spectral_radius  <-> pyamg/util/linalg.py:255-383 (Arnoldi estimate of rho)
"""
import numpy as np
from scipy import sparse


def get_diagonal(A, norm_eq=False, inv=False):
    """diag(A), diag(A^H A) (``norm_eq=1``) or diag(A A^H) (``norm_eq=2``), or its entry-wise inverse with 0 where
    the diagonal is 0 (pyamg/util/utils.py:530-600; the sums run over column-sorted rows, as there)."""
    if norm_eq in (1, 2):
        M = sparse.csr_array(A).copy()
        M.sort_indices()
        if norm_eq == 1:
            M = M.T
        D = np.asarray((M.multiply(M.conjugate())) @ np.ones((M.shape[0],)), dtype=np.float64)
    else:
        D = np.asarray(sparse.csr_array(A).diagonal(), dtype=np.float64)
    if inv:
        Dinv = np.zeros_like(D)
        mask = D != 0.0
        Dinv[mask] = 1.0 / D[mask]
        return Dinv
    return D


def get_block_diag(A, blocksize, inv_flag=True):
    """(n/bs, bs, bs) array of the diagonal blocks of A, pseudo-inverted if inv_flag."""
    if A.shape[0] != A.shape[1]:
        raise ValueError("Expected square matrix")
    if A.shape[0] % blocksize != 0:
        raise ValueError("blocksize and A.shape must be compatible")
    if not sparse.issparse(A) or A.format != "bsr" or A.blocksize != (blocksize, blocksize):
        A = sparse.bsr_array(A, blocksize=(blocksize, blocksize))
    nb = A.shape[0] // blocksize
    block_diag = np.zeros((nb, blocksize, blocksize), dtype=np.float64)
    rows = np.repeat(np.arange(nb), np.diff(A.indptr))
    on_diag = np.nonzero(rows == A.indices)[0]
    block_diag[A.indices[on_diag]] = A.data[on_diag]   # last duplicate wins
    if inv_flag:
        block_diag = np.linalg.pinv(block_diag)
    return np.ascontiguousarray(block_diag)


GPU_RHO_MIN_ROWS = 200_000


def gpu_rho_default(n):
    """Where the spectral-radius estimates run when the caller does not say: on the device for operators with at
    least GPU_RHO_MIN_ROWS rows when a CUDA device is visible (measured on a B200, profiles/r02_widening.jsonl:
    rho(D^-1 A) of the 2.1 M-row level-0 operator 0.17 s resident on the GPU vs 35 s on the host; values equal to
    3e-16), on the host otherwise.  AMGB_GPU_RHO=1 / 0 forces either."""
    import os
    env = os.environ.get("AMGB_GPU_RHO")
    if env in ("0", "1"):
        return env == "1"
    if n < GPU_RHO_MIN_ROWS:
        return False
    try:
        from . import _engine as E
        return E.lib().amgb_device_count() >= 1
    except Exception:                    # noqa: BLE001 - no library / no device: the host estimator
        return False


def _gpu_rho_enabled(where, n=0):
    if where is None:
        where = "gpu" if gpu_rho_default(n) else "host"
    if where not in ("host", "gpu"):
        raise ValueError("where must be 'host' or 'gpu'")
    return where == "gpu"


def _ritz_converged(H, ev, evec, k, m, tol):
    """linalg.py:351-365: |H[m, m-1] * (last component of the dominant Ritz vector)| / |ritz value| < tol."""
    if m >= H.shape[1] + 1 or abs(ev[k]) == 0.0:
        return False
    error = H[m, m - 1] * evec[-1, k]
    return bool(np.abs(error) / np.abs(ev[k]) < tol)


def _approximate_spectral_radius_gpu(A, row_scale, v0, maxiter, restarts, tol):
    """The rounds below on the device (amgb_arnoldi_*): the operator is uploaded once, every round
    of modified-Gram-Schmidt Arnoldi runs without a host round trip, only the small Hessenberg matrix comes back for
    the eigen-decomposition that picks the restart vector (a combination of the resident basis)."""
    import ctypes
    from . import _engine as E
    E.require_gpu()
    L = E.lib()
    keep = []
    M = E.as_matrix(sparse.csr_array(A), keep)
    sc = None if row_scale is None else np.ascontiguousarray(row_scale, dtype=np.float64)
    hdl = ctypes.c_void_p()
    import os
    E.check(L.amgb_arnoldi_create(int(os.environ.get("LOCAL_RANK", "0")), M, None if sc is None else E.f64p(sc), int(maxiter), ctypes.byref(hdl)))
    try:
        H = np.zeros((maxiter + 1, maxiter))
        m = ctypes.c_int32(0)
        rho = 0.0
        start = np.ascontiguousarray(v0, dtype=np.float64)
        for _ in range(restarts + 1):
            E.check(L.amgb_arnoldi_run(hdl, None if start is None else E.f64p(start), 1e-12, E.f64p(H.reshape(-1)),
                                       ctypes.byref(m)))
            if m.value == 0:
                break
            ev, evec = np.linalg.eig(H[:m.value, :m.value])
            k = int(np.argmax(np.abs(ev)))
            rho = float(np.abs(ev[k]))
            if _ritz_converged(H, ev, evec, k, m.value, tol) or m.value < maxiter:
                break
            coef = np.ascontiguousarray(np.real(evec[:, k]), dtype=np.float64)
            E.check(L.amgb_arnoldi_combine(hdl, E.f64p(coef), m.value))
            start = None
        return rho
    finally:
        L.amgb_arnoldi_destroy(hdl)


def _approximate_spectral_radius_host_native(A, row_scale, v0, maxiter, restarts, tol):
    """The same rounds with the multi-threaded host kernel (csrc/host_setup.cpp amgb_setup_arnoldi_round): the
    NumPy loop below spends its time in ~700 single-threaded vector passes per estimate, which dominates the SA
    setups of the large benchmark inputs."""
    from . import _host as Hh
    A = sparse.csr_array(A)
    n = A.shape[0]
    Ap = np.ascontiguousarray(A.indptr, dtype=np.int32)
    Aj = np.ascontiguousarray(A.indices, dtype=np.int32)
    Ax = np.ascontiguousarray(A.data, dtype=np.float64)
    sc = None if row_scale is None else np.ascontiguousarray(row_scale, dtype=np.float64)
    V = np.zeros((maxiter + 1, n))
    H = np.zeros((maxiter + 1, maxiter))
    V[0] = v0
    rho = 0.0
    for _ in range(restarts + 1):
        m = Hh.lib().amgb_setup_arnoldi_round(n, Hh.ip(Ap), Hh.ip(Aj), Hh.dp(Ax), None if sc is None else Hh.dp(sc),
                                              Hh.dp(V.reshape(-1)), maxiter, 1e-12, Hh.dp(H.reshape(-1)))
        if m == 0:
            break
        ev, evec = np.linalg.eig(H[:m, :m])
        k = int(np.argmax(np.abs(ev)))
        rho = float(np.abs(ev[k]))
        if _ritz_converged(H, ev, evec, k, m, tol) or m < maxiter:
            break
        V[0] = np.real(evec[:, k]) @ V[:m]                  # restart from the dominant Ritz vector
    return rho


def approximate_spectral_radius(A, maxiter=15, restarts=5, seed=20260922, row_scale=None, where=None, tol=0.01):
    """Largest |Ritz value| of a restarted Arnoldi process started from a seeded random vector.

    Same estimator as the reference (Arnoldi, 15 steps, up to 5 restarts, stop once the dominant Ritz pair's
    residual estimate is below ``tol`` = 0.01 relative; pyamg/util/linalg.py:255-383); the start vector is seeded
    here, so the value is reproducible (the reference's is not: SURVEY.md hazard 2).
    ``row_scale``: estimate rho(diag(row_scale) A) without forming the scaled matrix.  ``where='gpu'`` (or
    ``AMGB_GPU_RHO=1``) runs the Arnoldi rounds on the device -- same algorithm, values equal to rounding.
    """
    A = sparse.csr_array(A) if not sparse.issparse(A) else A
    n = A.shape[0]
    if n == 0:
        return 0.0
    rng = np.random.default_rng(seed)
    v0 = rng.random(n)
    maxiter = int(min(maxiter, n))
    if _gpu_rho_enabled(where, n) and sparse.issparse(A) and A.format == "csr":
        return _approximate_spectral_radius_gpu(A, row_scale, v0, maxiter, restarts, tol)
    if sparse.issparse(A) and A.format in ("csr", "bsr") and n >= 4096:
        return _approximate_spectral_radius_host_native(A, row_scale, v0, maxiter, restarts, tol)
    if row_scale is not None:
        A = sparse.dia_array((np.asarray(row_scale), 0), shape=(n, n)) @ sparse.csr_array(A)
    rho = 0.0
    for _ in range(restarts + 1):
        V = np.zeros((maxiter + 1, n))
        H = np.zeros((maxiter + 1, maxiter))
        nv = np.linalg.norm(v0)
        if nv == 0.0:
            break
        V[0] = v0 / nv
        m = maxiter
        for j in range(maxiter):
            w = A @ V[j]
            for i in range(j + 1):     # modified Gram-Schmidt
                H[i, j] = np.dot(V[i], w)
                w -= H[i, j] * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            if H[j + 1, j] < 1e-12 * max(1.0, abs(H[:j + 1, :j + 1]).max()):
                m = j + 1
                break
            V[j + 1] = w / H[j + 1, j]
        ev, evec = np.linalg.eig(H[:m, :m])
        k = int(np.argmax(np.abs(ev)))
        rho = float(np.abs(ev[k]))
        if _ritz_converged(H, ev, evec, k, m, tol) or m < maxiter:
            break
        v0 = np.real(V[:m].T @ evec[:, k])   # restart from the dominant Ritz vector
    return rho


def galerkin(R, A, P, where=None):
    """Coarse operator ``R @ A @ P`` (pyamg/classical/classical.py:201, aggregation/aggregation.py:425).

    ``where='gpu'`` (or ``AMGB_GPU_RAP=1``) computes the two products with the engine's SpGEMM
    (csrc/spgemm.cuh), which reproduces SciPy's ``csr_matmat`` bit for bit -- same values, same column order,
    same dropped zeros -- so the hierarchy is identical either way; the default stays SciPy on the host until the
    kernel has run on a B200.  CSR operands only."""
    import os
    if where is None:
        where = "gpu" if os.environ.get("AMGB_GPU_RAP") == "1" else "host"
    if where == "host":
        return R @ A @ P
    if where != "gpu":
        raise ValueError("galerkin: where must be 'host' or 'gpu'")
    from . import _engine as E
    return E.csr_matmat(E.csr_matmat(sparse.csr_array(R), sparse.csr_array(A)), sparse.csr_array(P))
