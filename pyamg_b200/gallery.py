"""Synthetic test operators for the BASELINE configs (host, NumPy/SciPy; not on the hot path).

Same matrices as the reference's gallery -- ``poisson`` (pyamg/gallery/laplacian.py:10),
``stencil_grid`` (pyamg/gallery/stencil.py:8-135: last grid dimension varies fastest, connections
across the boundary are dropped), ``diffusion_stencil_2d`` (pyamg/gallery/diffusion.py:15: rotated
anisotropic diffusion, Q1 finite elements or finite differences) -- so that bench inputs can be
generated on the GPU box, where the reference is not installed.  Written from the definitions
(index arithmetic over the stencil offsets), validated against the reference in
tests/golden/make_setup_golden.py.
"""
import numpy as np
from scipy import sparse


def stencil_grid(S, grid, dtype=np.float64, format="csr"):
    """Matrix of the constant-coefficient stencil S applied on a regular grid (Dirichlet cut-off)."""
    S = np.asarray(S, dtype=dtype)
    grid = tuple(int(g) for g in grid)
    if S.ndim != len(grid):
        raise ValueError("stencil dimension must equal number of grid dimensions")
    if not all(s % 2 == 1 for s in S.shape):
        raise ValueError("all stencil dimensions must be odd")
    if min(grid) < 1:
        raise ValueError("grid dimensions must be positive")
    N = int(np.prod(grid))
    strides = np.cumprod((1,) + grid[::-1])[:-1][::-1]          # last dimension fastest
    offs = [tuple(i - s // 2 for i, s in zip(idx, S.shape)) for idx in zip(*np.nonzero(S))]
    vals = [S[idx] for idx in zip(*np.nonzero(S))]
    shifts = [int(sum(o * st for o, st in zip(off, strides))) for off in offs]
    order = np.argsort(shifts, kind="stable")
    offs = [offs[k] for k in order]
    vals = np.array([vals[k] for k in order], dtype=dtype)
    shifts = np.array([shifts[k] for k in order], dtype=np.int64)
    # per-dimension validity of every offset (Dirichlet cut-off), combined by broadcasting
    ns = len(offs)
    valid = np.ones((N, ns), dtype=bool)
    for d, g in enumerate(grid):
        c = np.arange(g, dtype=np.int64)
        ok_d = np.stack([(c + off[d] >= 0) & (c + off[d] < g) for off in offs], axis=1)      # (g, ns)
        shape = [1] * len(grid) + [ns]
        shape[d] = g
        valid &= np.broadcast_to(ok_d.reshape(shape), grid + (ns,)).reshape(N, ns)
    if len(set(shifts.tolist())) == ns and N < 2**31 - 1:
        # fast path: offsets sorted by shift == columns ascending inside every row -> CSR directly
        indptr = np.zeros(N + 1, dtype=np.int64)
        np.cumsum(valid.sum(axis=1), out=indptr[1:])
        cols = (np.arange(N, dtype=np.int32)[:, None] + shifts.astype(np.int32)[None, :])[valid]
        data = np.broadcast_to(vals[None, :], (N, ns))[valid]
        A = sparse.csr_array((data, cols, indptr.astype(np.int32)), shape=(N, N))
    else:
        # degenerate grids (a dimension shorter than the stencil) alias offsets: sum duplicates
        r, k = np.nonzero(valid)
        A = sparse.coo_array((vals[k], (r, r + shifts[k])), shape=(N, N)).tocsr()
        A.indptr = A.indptr.astype(np.int32)
        A.indices = A.indices.astype(np.int32)
    return A.asformat(format)


def poisson(grid, dtype=np.float64, format="csr"):
    """N-d Poisson operator on a regular grid: the [-1 2 -1] stencil summed over the dimensions."""
    grid = tuple(grid)
    nd = len(grid)
    S = np.zeros((3,) * nd, dtype=dtype)
    c = (1,) * nd
    for d in range(nd):
        for o in (0, 2):
            i = list(c)
            i[d] = o
            S[tuple(i)] = -1
    S[c] = 2 * nd
    return stencil_grid(S, grid, dtype=dtype, format=format)


def diffusion_stencil_2d(epsilon=1.0, theta=0.0, type="FE"):
    """3x3 stencil of -div Q diag(1, eps) Q^T grad u, rotation angle theta (y varies first)."""
    eps, theta = float(epsilon), float(theta)
    C, S = np.cos(theta), np.sin(theta)
    CS, CC, SS = C * S, C * C, S * S
    if type == "FE":     # bilinear (Q1) elements, weak form (K grad u, grad v) on the unit square mesh
        a = (-eps - 1) * (CC + SS) + (3 * eps - 3) * CS
        b = (2 * eps - 4) * CC + (-4 * eps + 2) * SS
        c = (-eps - 1) * (CC + SS) + (-3 * eps + 3) * CS
        d = (-4 * eps + 2) * CC + (2 * eps - 4) * SS
        e = (8 * eps + 8) * (CC + SS)
        return np.array([[a, b, c], [d, e, d], [c, b, a]]) / 6.0
    if type == "FD":     # central differences, h = 1
        Exx = CC + eps * SS
        Fxy = 2 * (1 - eps) * CS
        Gyy = eps * CC + SS
        return np.array([[-Fxy / 4, -Exx, Fxy / 4], [-Gyy, 2 * Exx + 2 * Gyy, -Gyy], [Fxy / 4, -Exx, -Fxy / 4]])
    raise ValueError("type must be 'FE' or 'FD'")
