"""CPU restatement of the reference's Householder GMRES and flexible GMRES.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py): the checker of `amgb_solve_gmres`, never imported by the product.

Follows, statement by statement,
  pyamg/krylov/_gmres_householder.py:21-360  (what ``pyamg.krylov.gmres`` resolves to by default,
                                              _gmres.py:11-127 ``orthog='householder'``; LEFT preconditioning), and
  pyamg/krylov/_fgmres.py:17-345             (RIGHT preconditioning, preconditioned vectors stored in Z),
with the native helpers of pyamg/amg_core/krylov.h restated in NumPy:
  apply_householders :37-62, householder_hornerscheme :106-135, apply_givens :158-187
and the norm of pyamg/util/linalg.py:13-55 (sqrt of np.inner).  ``ml.solve(accel='gmres' | 'fgmres')`` reaches
these through multilevel.py:479-508 with ``M = aspreconditioner(cycle)`` (one cycle from a zero guess).

Pinned by tests/test_oracle.py against goldens the real reference produced (``x_ref_gmres``, ``x_ref_fgmres``,
residual histories, info flags).
"""
import numpy as np
import scipy.linalg
from scipy.linalg import get_lapack_funcs


def _norm(x):
    x = np.ravel(x)
    return np.sqrt(np.inner(x.conj(), x).real)


def _mysign(x):
    if x == 0.0:
        return 1.0
    return x / np.abs(x)


def _apply_householders(z, W, start, stop, step):
    """krylov.h:37-62: z <- (I - 2 w_j w_j^T) z for j = start, start+step, ... (stop exclusive)."""
    for j in range(start, stop, step):
        alpha = np.dot(W[j, :], z)
        alpha *= -2
        z += alpha * W[j, :]


def _householder_hornerscheme(z, W, y, start, stop, step):
    """krylov.h:106-135."""
    for j in range(start, stop, step):
        z[j] += y[j]
        alpha = np.dot(W[j, :], z)
        alpha *= -2
        z += alpha * W[j, :]


def _apply_givens(Q, v, nrot):
    """krylov.h:158-187."""
    for rot in range(nrot):
        t = v[rot]
        v[rot] = Q[4 * rot] * t + Q[4 * rot + 1] * v[rot + 1]
        v[rot + 1] = Q[4 * rot + 2] * t + Q[4 * rot + 3] * v[rot + 1]


def _iteration_limits(n, restart, maxiter):
    """_gmres_householder.py:129-149 / _fgmres.py:139-159."""
    if restart:
        max_outer = maxiter if maxiter else 1
        restart = min(restart, n)
        return max_outer, restart
    if maxiter is None:
        maxiter = min(n, 40)
    elif maxiter > n:
        maxiter = n
    return 1, maxiter


def gmres(A, b, x0=None, tol=1e-5, restart=None, maxiter=None, M=None, callback=None, residuals=None,
          flexible=False, matvec=None):
    """``flexible=False``: gmres_householder; ``True``: fgmres.  ``M`` and ``matvec`` are callables (v -> M v, A v);
    by default ``matvec`` is ``A @ v``."""
    if matvec is None:
        def matvec(v):
            return A @ v
    if M is None:
        def M(v):
            return v.copy()
    b = np.ravel(np.asarray(b, dtype=np.float64))
    n = len(b)
    x = np.zeros(n) if x0 is None else np.array(np.ravel(x0), dtype=np.float64)
    [lartg] = get_lapack_funcs(["lartg"], [x])
    max_outer, max_inner = _iteration_limits(n, restart, maxiter)
    if n == 1:
        entry = np.ravel(matvec(np.array([1.0])))
        return b / entry, 0

    r = b - matvec(x)
    if not flexible:
        r = M(r)
    normr = _norm(r)
    if residuals is not None:
        residuals[:] = [normr]
    normb = _norm(b)
    if normb == 0.0:
        scale = 1.0
    else:
        scale = normb if flexible else _norm(M(b))        # fgmres: ||b||; gmres: ||M b||
    if normr < tol * scale:
        return x, 0
    niter = 0

    for _outer in range(max_outer):
        w = r
        beta = _mysign(w[0]) * normr
        w[0] = w[0] + beta
        w[:] = w / _norm(w)
        Q = np.zeros(4 * max_inner)
        H = np.zeros((max_inner, max_inner))
        W = np.zeros((max_inner if flexible else max_inner + 1, n))
        if flexible:
            Z = np.zeros((n, max_inner))
        W[0, :] = w
        g = np.zeros(n)
        g[0] = -beta

        for inner in range(max_inner):
            v = -2.0 * np.conjugate(w[inner]) * w
            v[inner] = v[inner] + 1.0
            _apply_householders(v, W, inner - 1, -1, -1)
            if flexible:
                v = M(v)
                Z[:, inner] = v
                v = matvec(v)
            else:
                v = matvec(v)
                v = M(v)
            v = np.array(np.ravel(v), dtype=np.float64)
            _apply_householders(v, W, 0, inner + 1, 1)

            if inner != n - 1:
                if inner < (max_inner - 1):
                    w = W[inner + 1, :]
                vslice = v[inner + 1:]
                alpha = _norm(vslice)
                if alpha != 0:
                    alpha = _mysign(vslice[0]) * alpha
                    if inner < (max_inner - 1):
                        w[inner + 1:] = vslice
                        w[inner + 1] += alpha
                        w[:] = w / _norm(w)
                    v[inner + 1] = -alpha
                    v[inner + 2:] = 0.0

            if inner > 0:
                _apply_givens(Q, v, inner)

            if inner != n - 1:
                if v[inner + 1] != 0:
                    c, s, _r = lartg(v[inner], v[inner + 1])
                    Qblock = np.array([[c, s], [-np.conjugate(s), c]])
                    Q[(inner * 4):((inner + 1) * 4)] = np.ravel(Qblock).copy()
                    g[inner:inner + 2] = np.dot(Qblock, g[inner:inner + 2])
                    v[inner] = np.dot(Qblock[0, :], v[inner:inner + 2])
                    v[inner + 1] = 0.0

            H[:, inner] = v[0:max_inner]
            if not flexible:
                niter += 1

            if inner < max_inner - 1:
                normr = np.abs(g[inner + 1])
                if normr < tol * scale:
                    if flexible:
                        pass                      # _fgmres.py:296-297: break happens BEFORE niter += 1
                    break
                if residuals is not None:
                    residuals.append(normr)
                if callback is not None:
                    y = scipy.linalg.solve(H[0:(inner + 1), 0:(inner + 1)], g[0:(inner + 1)])
                    if flexible:
                        update = np.dot(Z[:, 0:inner + 1], y)
                    else:
                        update = np.zeros(n)
                        _householder_hornerscheme(update, W, y, inner, -1, -1)
                    callback(x + update)
            if flexible:
                niter += 1

        y = scipy.linalg.solve(H[0:(inner + 1), 0:(inner + 1)], g[0:(inner + 1)])
        if flexible:
            update = np.dot(Z[:, 0:inner + 1], y)
            x = x + update
        else:
            update = np.zeros(n)
            _householder_hornerscheme(update, W, y, inner, -1, -1)
            x[:] = x + update
        r = b - matvec(x)
        if not flexible:
            r = M(r)
        normr = _norm(r)
        if callback is not None:
            callback(x)
        if residuals is not None:
            residuals.append(normr)
        indices = x != 0
        if indices.any():
            change = np.max(np.abs(update[indices] / x[indices]))
            if change < 1e-12:
                return x, -1
        if normr < tol * scale:
            return x, 0

    return x, niter


def bicgstab(A, b, x0=None, tol=1e-5, maxiter=None, M=None, callback=None, residuals=None, matvec=None):
    """pyamg/krylov/_bicgstab.py:10-200 (criteria 'rr'): right-preconditioned BiCGStab, statement by statement."""
    if matvec is None:
        def matvec(v):
            return A @ v
    if M is None:
        def M(v):
            return v.copy()
    b = np.ravel(np.asarray(b, dtype=np.float64))
    n = len(b)
    x = np.zeros(n) if x0 is None else np.array(np.ravel(x0), dtype=np.float64)
    if maxiter is None:
        maxiter = len(x) + 5
    elif maxiter < 1:
        raise ValueError("Number of iterations must be positive")
    r = b - matvec(x)
    normr = _norm(r)
    if residuals is not None:
        residuals[:] = [normr]
    normb = _norm(b)
    if normb == 0.0:
        normb = 1.0
    rtol = tol * normb
    if normr < rtol:
        return x, 0
    if n == 1:
        entry = np.ravel(matvec(np.array([1.0])))
        return b / entry, 0
    rstar = r.copy()
    p = r.copy()
    rrstarOld = np.inner(rstar.conjugate(), r)
    it = 0
    while True:
        Mp = M(p)
        AMp = matvec(Mp)
        alpha = rrstarOld / np.inner(rstar.conjugate(), AMp)
        s = r - alpha * AMp
        Ms = M(s)
        AMs = matvec(Ms)
        omega = np.inner(AMs.conjugate(), s) / np.inner(AMs.conjugate(), AMs)
        x = x + alpha * Mp + omega * Ms
        r = s - omega * AMs
        rrstarNew = np.inner(rstar.conjugate(), r)
        beta = (rrstarNew / rrstarOld) * (alpha / omega)
        rrstarOld = rrstarNew
        p = r + beta * (p - omega * AMp)
        it += 1
        normr = _norm(r)
        if residuals is not None:
            residuals.append(normr)
        if callback is not None:
            callback(x)
        if normr < rtol:
            return x, 0
        if it == maxiter:
            return x, it
