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


def _q1_plane_strain_stiffness(dx, dy, lame, mu):
    """8x8 stiffness of one bilinear (Q1) rectangle dx x dy in plane strain, by 2x2 Gauss quadrature (exact for
    this element).  Vertices counter-clockwise from the lower left; unknowns (u_x, u_y) per vertex."""
    D = np.array([[lame + 2.0 * mu, lame, 0.0], [lame, lame + 2.0 * mu, 0.0], [0.0, 0.0, mu]])
    sx = np.array([-1.0, 1.0, 1.0, -1.0])        # reference-square signs of the four vertices
    sy = np.array([-1.0, -1.0, 1.0, 1.0])
    g = 1.0 / np.sqrt(3.0)
    K = np.zeros((8, 8))
    for xi in (-g, g):
        for eta in (-g, g):
            dNdx = sx * (1.0 + sy * eta) / 4.0 * (2.0 / dx)
            dNdy = sy * (1.0 + sx * xi) / 4.0 * (2.0 / dy)
            Bm = np.zeros((3, 8))
            Bm[0, 0::2] = dNdx
            Bm[1, 1::2] = dNdy
            Bm[2, 0::2] = dNdy
            Bm[2, 1::2] = dNdx
            K += Bm.T @ D @ Bm * (dx * dy / 4.0)
    return K


def linear_elasticity(grid, spacing=None, E=1e5, nu=0.3, format=None):
    """2-D linear elasticity, Q1 elements on a regular grid with a clamped (Dirichlet) boundary:
    (A, B) = BSR(2,2) stiffness matrix over the grid[0] x grid[1] interior vertices and the three rigid-body
    modes (two translations, one rotation about the grid centre) as near-nullspace candidates.

    Same problem as pyamg.gallery.linear_elasticity (gallery/elasticity.py:9-120; vertex k = iy*X + ix, unknowns
    interleaved); the element matrix is integrated here by quadrature instead of the reference's closed form, so
    entries agree to rounding (tests/test_setup.py).  The reference only accepts square grids (its index
    arithmetic overruns otherwise); any X x Y grid is assembled here."""
    if len(grid) != 2:
        raise NotImplementedError(f"No support for grid={grid}")
    X, Y = int(grid[0]), int(grid[1])
    if X < 1 or Y < 1:
        raise ValueError("invalid grid shape")
    DX, DY = (1.0, 1.0) if spacing is None else (float(spacing[0]), float(spacing[1]))
    lame = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
    mu = E / (2.0 + 2.0 * nu)
    K = _q1_plane_strain_stiffness(DX, DY, lame, mu)
    # elements (ex, ey), 0 <= ex <= X, 0 <= ey <= Y, on the (X+2) x (Y+2) vertex grid; interior vertices only
    ex, ey = np.meshgrid(np.arange(X + 1), np.arange(Y + 1), indexing="xy")
    ex, ey = ex.ravel(), ey.ravel()
    vx = np.stack([ex, ex + 1, ex + 1, ex], axis=1)          # vertex grid coordinates, counter-clockwise
    vy = np.stack([ey, ey, ey + 1, ey + 1], axis=1)
    inside = (vx >= 1) & (vx <= X) & (vy >= 1) & (vy <= Y)
    vid = (vy - 1) * X + (vx - 1)                              # interior vertex number
    dof = np.stack([2 * vid, 2 * vid + 1], axis=2).reshape(-1, 8)
    ok = np.repeat(inside, 2, axis=1)
    I = np.repeat(dof[:, :, None], 8, axis=2)
    J = np.repeat(dof[:, None, :], 8, axis=1)
    M = ok[:, :, None] & ok[:, None, :]
    V = np.broadcast_to(K, I.shape)
    n = 2 * X * Y
    A = sparse.coo_array((V[M], (I[M], J[M])), shape=(n, n)).tocsr().tobsr(blocksize=(2, 2))
    # rigid-body modes at the interior vertices; coordinates relative to the centre of the clamped plate
    iy, ix = np.divmod(np.arange(X * Y), X)
    px = (ix + 1 - (X + 1) / 2.0) * DX
    py = (iy + 1 - (Y + 1) / 2.0) * DY
    B = np.zeros((n, 3))
    B[0::2, 0] = 1.0
    B[1::2, 1] = 1.0
    B[0::2, 2] = -py
    B[1::2, 2] = px
    return (A if format is None else A.asformat(format)), B
