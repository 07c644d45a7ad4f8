"""CPU emulation of the flat-gather tile kernel's per-tile algorithm (csrc/tile_flat_kernel.cuh, opt-in
AMGB_TILE_FLAT=1): the staging offsets, the in-place products, the diagonal test on the staged row pointers and
the per-tile lane grouping -- lane by lane in numpy, on tiles from the engine's own tile builder -- against the
oracle's sweeps.  It checks the LOGIC the CUDA code encodes (not CUDA semantics; those need the B200:
tests/test_gpu_experimental.py)."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from pyamg_b200 import _engine as E
from pyamg_b200.gallery import poisson

T, RMAX = 224, 64
OP_SPMV, OP_RESID, OP_PADD, OP_JACOBI, OP_GS = range(5)


def build_tiles(Ap, breaks=None):
    n = len(Ap) - 1
    cap = n + 2 + (0 if breaks is None else len(breaks))
    row0, nz0 = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    nt = ctypes.c_int32(0)
    br = None if breaks is None else np.asarray(breaks, np.int64)
    tp = None if breaks is None else np.zeros(len(br), np.int32)
    E.check(E.lib().amgb_debug_build_tiles(n, E.i32p(np.ascontiguousarray(Ap, np.int32)), 32,
                                           br.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) if br is not None else None,
                                           0 if br is None else len(br) - 1, T, RMAX, E.i32p(row0), E.i32p(nz0), cap,
                                           E.i32p(tp) if tp is not None else None, ctypes.byref(nt)))
    return row0[:nt.value + 1], nz0[:nt.value + 1], tp


def flat_tile(op, t, row0s, nz0s, Ap, Aj, Ax, x, b, y, omega, r=None):
    """One tile, exactly as csr_tile_flat_kernel walks it (32 lanes emulated sequentially, phase by phase)."""
    r0, r1 = int(row0s[t]), int(row0s[t + 1])
    s0, s1 = int(nz0s[t]), int(nz0s[t + 1])
    nrows, ln = r1 - r0, s1 - s0
    need_diag = op in (OP_JACOBI, OP_GS)
    assert ln <= T, "long-row tiles take the other branch (copied from the validated kernel)"
    # --- what the TMA stages (tile_issue): widened, 16-byte aligned segments
    s4 = s0 & ~3
    cnt = ((s1 + 3) & ~3) - s4
    r4 = r0 & ~3
    rcnt = ((r1 + 1 + 3) & ~3) - r4
    v2 = r0 & ~1
    vcnt = ((r1 + 1) & ~1) - v2
    pad = lambda a, lo, k: np.concatenate([a[lo:lo + k], np.zeros(max(0, lo + k - len(a)), a.dtype)])
    val = pad(Ax, s4, cnt).astype(np.float64)
    col = pad(Aj, s4, cnt)
    ptr = pad(Ap, r4, rcnt)
    bseg = pad(b, v2, vcnt) if b is not None else None
    xseg = pad(x, v2, vcnt)
    soff, poff, voff = s4, r4, v2
    e0, e1 = s0 - soff, s0 - soff + ln
    dslot = np.zeros(RMAX)
    # --- phase 1
    EPL = T // 32
    for lane in range(32):
        for u in range(EPL):
            e = e0 + lane + 32 * u
            if e >= e1:
                continue
            c = int(col[e])
            v = val[e]
            p = v * x[c]
            if need_diag and r0 <= c < r0 + nrows:
                ge = e + soff
                lr = c - r0
                if ptr[r0 + lr - poff] <= ge < ptr[r0 + lr + 1 - poff]:
                    dslot[lr] = v
                    p = 0.0
            val[e] = p
    # --- phase 2
    g = 32
    while g > 1 and g * nrows > 32:
        g >>= 1
    rpp = 32 // g
    r2 = 0.0
    for rbase in range(0, nrows, rpp):
        sums = np.zeros(32)
        for lane in range(32):
            sub, grp = lane & (g - 1), lane // g
            lr = rbase + grp
            if lr < nrows:
                row = r0 + lr
                jb, je = ptr[row - poff] - soff, ptr[row - poff + 1] - soff
                for jj in range(jb + sub, je, g):
                    sums[lane] += val[jj]
        o = g >> 1
        while o > 0:
            sums = sums + sums[np.arange(32) ^ o]
            o >>= 1
        for lane in range(0, 32, g):
            lr = rbase + lane // g
            if lr >= nrows:
                continue
            row, s = r0 + lr, sums[lane]
            if op == OP_SPMV:
                y[row] = s
            elif op == OP_RESID:
                y[row] = bseg[row - voff] - s
                r2 += y[row] ** 2
            elif op == OP_PADD:
                y[row] += s
            elif op == OP_JACOBI:
                xi, bi, d = xseg[row - voff], bseg[row - voff], dslot[lr]
                y[row] = (1.0 - omega) * xi + omega * ((bi - s) / d) if d != 0.0 else xi
                if r is not None:
                    r[row] = bi - s - d * xi
            else:
                d = dslot[lr]
                if d != 0.0:
                    gs = (bseg[row - voff] - s) / d
                    y[row] = gs if omega == 1.0 else omega * gs + (1.0 - omega) * y[row]
    return r2


def matrices():
    rng = np.random.default_rng(3)
    A = poisson((9, 8, 7)).tocsr()
    yield "poisson3d", A
    B = (A @ A).tocsr()                       # ~25 entries per row
    B.sort_indices()
    yield "poisson3d_squared", B
    n = 150
    M = sp.random(n, n, density=0.3, random_state=5, format="csr") + sp.eye(n) * 7.0
    M = M.tolil()
    M[3, 3] = 0.0                             # a row without a stored diagonal
    M[40, :] = 0.0                            # an empty row
    M = M.tocsr()
    M.eliminate_zeros()
    yield "random_dense_rows", M
    lens = rng.integers(0, 40, 90)
    rows = np.repeat(np.arange(90), lens)
    cols = rng.integers(0, 90, len(rows))
    R = sp.csr_matrix((rng.random(len(rows)), (rows, cols)), shape=(90, 90))   # sums duplicates
    yield "ragged", (R + sp.eye(90) * 3).tocsr()


@pytest.mark.parametrize("name,A", list(matrices()), ids=[m[0] for m in matrices()])
def test_flat_tile_algorithm_matches_oracle(name, A):
    A = A.tocsr()
    n = A.shape[0]
    Ap, Aj, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    rng = np.random.default_rng(11)
    x, b = rng.random(n), rng.random(n)
    row0s, nz0s, _ = build_tiles(Ap)
    nt = len(row0s) - 1
    assert np.all(np.diff(nz0s) <= T)
    # SpMV / residual / prolong-add
    y = np.zeros(n)
    for t in range(nt):
        flat_tile(OP_SPMV, t, row0s, nz0s, Ap, Aj, Ax, x, None, y, 0.0)
    assert np.allclose(y, A @ x, rtol=1e-14, atol=1e-14)
    y = np.zeros(n)
    r2 = sum(flat_tile(OP_RESID, t, row0s, nz0s, Ap, Aj, Ax, x, b, y, 0.0) for t in range(nt))
    assert np.allclose(y, b - A @ x, rtol=1e-13, atol=1e-13) and np.isclose(r2, np.sum((b - A @ x) ** 2))
    y = b.copy()
    for t in range(nt):
        flat_tile(OP_PADD, t, row0s, nz0s, Ap, Aj, Ax, x, None, y, 0.0)
    assert np.allclose(y, b + A @ x, rtol=1e-14, atol=1e-14)
    # weighted Jacobi with the residual by-product (relaxation.h:309-346)
    y, r = np.zeros(n), np.zeros(n)
    for t in range(nt):
        flat_tile(OP_JACOBI, t, row0s, nz0s, Ap, Aj, Ax, x, b, y, 0.7, r)
    xo = x.copy()
    oracle.jacobi(A, xo, b, iterations=1, omega=0.7)
    assert np.allclose(y, xo, rtol=1e-13, atol=1e-14)
    assert np.allclose(r, b - A @ x, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name,A", list(matrices())[:3], ids=[m[0] for m in list(matrices())[:3]])
@pytest.mark.parametrize("omega", [1.0, 1.3])
def test_flat_tile_gauss_seidel_waves_match_oracle(name, A, omega):
    """Wave-major permuted operator, tiles that never cross a wave boundary, one launch per wave."""
    A = A.tocsr()
    A.sort_indices()
    n = A.shape[0]
    Ap, Aj = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    wave_of = np.zeros(n, np.int32)
    nw = ctypes.c_int32(0)
    E.check(E.lib().amgb_wave_schedule(n, E.i32p(Ap), E.i32p(Aj), None, n, E.i32p(wave_of), ctypes.byref(nw)))
    order = np.argsort(wave_of, kind="stable")
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)
    B = A[order][:, order].tocsr()              # columns re-sorted: only the summation order changes
    Bp, Bj, Bx = B.indptr.astype(np.int32), B.indices.astype(np.int32), B.data.astype(np.float64)
    breaks = np.concatenate([[0], np.cumsum(np.bincount(wave_of, minlength=nw.value + 1)[1:])])
    row0s, nz0s, tp = build_tiles(Bp, breaks)
    rng = np.random.default_rng(12)
    x, b = rng.random(n), rng.random(n)
    xp, bp = x[order].copy(), b[order].copy()
    for w in range(nw.value):
        for t in range(tp[w], tp[w + 1]):
            # y aliases x: the wave's rows only read columns of other waves (and their own diagonal)
            flat_tile(OP_GS, t, row0s, nz0s, Bp, Bj, Bx, xp, bp, xp, omega)
    xo = x.copy()
    if omega == 1.0:
        oracle.gauss_seidel(A, xo, b, iterations=1, sweep="forward")
    else:
        oracle.sor(A, xo, b, omega, iterations=1, sweep="forward")
    assert np.allclose(xp[pos], xo, rtol=1e-12, atol=1e-13)
