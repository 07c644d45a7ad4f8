"""Known-answer vectors of the reference's own smoother tests, restated once and run against BOTH
the CPU oracle (tests/test_oracle.py) and the CUDA kernels (tests/test_gpu_relaxation.py).

Source: pyamg/relaxation/tests/test_relaxation.py -- test_jacobi :148-197, test_gauss_seidel_csr
:299-362, test_gauss_seidel_indexed :364-411.  The matrix is the 1-D Poisson stencil [-1 2 -1].
Each entry: (routine, N, x0, b, kwargs, expected).
"""
import numpy as np
import scipy.sparse as sp


def poisson1d(n):
    return sp.diags_array([2 * np.ones(n), -np.ones(n), -np.ones(n)], offsets=[0, -1, 1],
                          shape=(n, n), format="csr")


def _a(*v):
    return np.array(v, dtype=np.float64)


J3 = _a(5.5, 11.0, 15.5)
KATS = [
    # test_jacobi :148-197
    ("jacobi", 1, _a(0), _a(0), {}, _a(0)),
    ("jacobi", 3, _a(0, 0, 0), _a(0, 1, 2), {}, _a(0.0, 0.5, 1.0)),
    ("jacobi", 3, _a(0, 1, 2), _a(0, 0, 0), {}, _a(0.5, 1.0, 0.5)),
    ("jacobi", 1, _a(0), _a(10), {}, _a(5)),
    ("jacobi", 3, _a(0, 1, 2), _a(10, 20, 30), {}, J3),
    ("jacobi", 3, _a(0, 1, 2), _a(10, 20, 30), {"omega": 1.0 / 3.0}, 2.0 / 3.0 * _a(0, 1, 2) + 1.0 / 3.0 * J3),
    # test_gauss_seidel_csr :299-346
    ("gauss_seidel", 1, _a(0), _a(0), {}, _a(0)),
    ("gauss_seidel", 3, _a(0, 1, 2), _a(0, 0, 0), {}, _a(1.0 / 2.0, 5.0 / 4.0, 5.0 / 8.0)),
    ("gauss_seidel", 1, _a(0), _a(0), {"sweep": "backward"}, _a(0)),
    ("gauss_seidel", 3, _a(0, 1, 2), _a(0, 0, 0), {"sweep": "backward"}, _a(1.0 / 8.0, 1.0 / 4.0, 1.0 / 2.0)),
    ("gauss_seidel", 1, _a(0), _a(10), {}, _a(5)),
    ("gauss_seidel", 3, _a(0, 1, 2), _a(10, 20, 30), {}, _a(11.0 / 2.0, 55.0 / 4, 175.0 / 8.0)),
    # test_gauss_seidel_indexed :364-411
    ("gauss_seidel_indexed", 1, _a(0), _a(0), {"indices": [0]}, _a(0)),
    ("gauss_seidel_indexed", 3, _a(0, 1, 2), _a(0, 0, 0), {"indices": [0, 1, 2]}, _a(1.0 / 2.0, 5.0 / 4.0, 5.0 / 8.0)),
    ("gauss_seidel_indexed", 3, _a(0, 1, 2), _a(0, 0, 0), {"indices": [2, 1, 0], "sweep": "backward"},
     _a(1.0 / 2.0, 5.0 / 4.0, 5.0 / 8.0)),
    ("gauss_seidel_indexed", 3, _a(0, 1, 2), _a(0, 0, 0), {"indices": [0, 1, 2], "sweep": "backward"},
     _a(1.0 / 8.0, 1.0 / 4.0, 1.0 / 2.0)),
    ("gauss_seidel_indexed", 4, _a(1, 1, 1, 1), _a(0, 0, 0, 0), {"indices": [0, 3]}, _a(0.5, 1.0, 1.0, 0.5)),
    ("gauss_seidel_indexed", 4, _a(1, 1, 1, 1), _a(0, 0, 0, 0), {"indices": [0, 0]}, _a(0.5, 1.0, 1.0, 1.0)),
]


def run_kats(mod):
    """Run every KAT through `mod.<routine>(A, x, b, **kw)`; returns list of (case, got, expected)."""
    out = []
    for fn, n, x0, b, kw, exp in KATS:
        x = x0.copy()
        getattr(mod, fn)(poisson1d(n), x, b.copy(), **kw)
        out.append(((fn, n, kw), x, exp))
    return out


def gs_convergence_case(mod):
    """test_relaxation.py:348-362: forward and backward GS converge alike on x=1, b=0, N=100."""
    A = poisson1d(100)
    b = np.zeros(100)
    x = np.ones(100)
    mod.gauss_seidel(A, x, b, iterations=200, sweep="forward")
    r1 = np.linalg.norm(A @ x, 2)
    x = np.ones(100)
    mod.gauss_seidel(A, x, b, iterations=200, sweep="backward")
    r2 = np.linalg.norm(A @ x, 2)
    return r1, r2


def dense_gs_gold(A, x, b, sweep):
    """test_relaxation.py:251-297: Gauss-Seidel as dense triangular solves."""
    D, Lo, U = np.diag(np.diag(A)), np.tril(A, -1), np.triu(A, 1)
    g = x.copy()
    if sweep in ("forward", "symmetric"):
        g = np.linalg.solve(D + Lo, b - U @ g)
    if sweep in ("backward", "symmetric"):
        g = np.linalg.solve(D + U, b - Lo @ g)
    return g


def block_jacobi_gold(A, x, b, Dinv, bs, omega):
    """test_relaxation.py:1517-1777 style dense gold for block Jacobi."""
    g = x.copy()
    for i in range(A.shape[0] // bs):
        sl = slice(i * bs, (i + 1) * bs)
        r = b[sl] - A[sl] @ x + A[sl, sl] @ x[sl]
        g[sl] = (1 - omega) * x[sl] + omega * (Dinv[i] @ r)
    return g
