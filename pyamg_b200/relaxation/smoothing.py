"""Smoother factory: resolve a smoother spec into per-level closures.

Mirror of pyamg/relaxation/smoothing.py for the smoothers on the hot path:
``change_smoothers`` :75-369 (spec grammar ``name | (name, {opts}) | [per-level list] | None``,
short lists repeat their last entry, ``ml.symmetric_smoothing`` truth table), ``setup_jacobi``
:501-508, ``setup_gauss_seidel`` :494-498, ``setup_block_jacobi`` :552-579, ``setup_sor`` :620-624,
``setup_none`` :833-837, ``rebuild_smoother`` :881-908.  Extra (not in the reference registry):
``gauss_seidel_indexed`` / ``multicolor_gauss_seidel`` which install the reference's
``relaxation.gauss_seidel_indexed`` over a colour-sorted row list -- BASELINE config 3.

What is stored on the level is exactly what the reference stores: a ``functools.partial`` of the
relaxation function whose ``__name__`` is the registry key and whose ``.keywords`` carry
``iterations``, ``omega`` (already divided by rho), ``sweep``, ``Dinv``, ``blocksize``.  The cycle
engine parses those closures (``describe``), so hierarchies built by the reference and
hierarchies equipped here are interchangeable.
"""
from functools import partial, update_wrapper

import numpy as np
from scipy import sparse

from . import relaxation
from .chebyshev import chebyshev_polynomial_coefficients
from ..util import approximate_spectral_radius, get_block_diag, get_diagonal
from .. import _engine as E

DEFAULT_SWEEP = "forward"
DEFAULT_NITER = 1
SYMMETRIC_RELAXATION = ["jacobi", "richardson", "block_jacobi", "jacobi_ne", "chebyshev", None]   # smoothing.py:49-50


def _unpack_arg(v):
    if isinstance(v, tuple):
        return v[0], v[1]
    return v, {}


def rho_D_inv_A(A):
    """(approx.) spectral radius of D^-1 A, cached on the matrix (smoothing.py:372-400)."""
    if not hasattr(A, "rho_D_inv"):
        from ..util import gpu_rho_default
        D_inv = get_diagonal(A, inv=True)
        if A.format == "csr" and gpu_rho_default(A.shape[0]):
            # Arnoldi rounds on the device, D^-1 applied as a row scaling (no scaled copy of A is formed)
            A.rho_D_inv = approximate_spectral_radius(A, row_scale=D_inv, where="gpu")
        else:
            D_inv_A = sparse.dia_array((D_inv, 0), shape=(len(D_inv), len(D_inv))) @ sparse.csr_array(A)
            A.rho_D_inv = approximate_spectral_radius(D_inv_A, where="host")
    return A.rho_D_inv


def rho_block_D_inv_A(A, Dinv):
    """(approx.) spectral radius of blockdiag(A)^-1 A (smoothing.py:403-449)."""
    if not hasattr(A, "rho_block_D_inv"):
        bs = Dinv.shape[1]
        nb = Dinv.shape[0]
        Dinv_bsr = sparse.bsr_array((Dinv, np.arange(nb), np.arange(nb + 1)), shape=A.shape)
        A.rho_block_D_inv = approximate_spectral_radius(Dinv_bsr @ sparse.bsr_array(A, blocksize=(bs, bs)))
    return A.rho_block_D_inv


def setup_gauss_seidel(lvl, iterations=DEFAULT_NITER, sweep=DEFAULT_SWEEP):
    smoother = partial(relaxation.gauss_seidel, iterations=iterations, sweep=sweep)
    update_wrapper(smoother, relaxation.gauss_seidel)
    return smoother


def setup_jacobi(lvl, iterations=DEFAULT_NITER, omega=1.0, withrho=True):
    if withrho:
        omega = omega / rho_D_inv_A(lvl.A)
    smoother = partial(relaxation.jacobi, iterations=iterations, omega=omega)
    update_wrapper(smoother, relaxation.jacobi)
    return smoother


def setup_block_jacobi(lvl, iterations=DEFAULT_NITER, omega=1.0, Dinv=None, blocksize=None,
                       withrho=True):
    if blocksize is None and Dinv is None:
        if sparse.issparse(lvl.A) and lvl.A.format == "csr":
            blocksize = 1
        elif sparse.issparse(lvl.A) and lvl.A.format == "bsr":
            blocksize = lvl.A.blocksize[0]
    elif blocksize is None:
        blocksize = Dinv.shape[1]
    if blocksize == 1:
        smoother = setup_jacobi(lvl, iterations=iterations, omega=omega, withrho=withrho)
        update_wrapper(smoother, relaxation.block_jacobi)   # __name__ stays the registry key
        return smoother
    if Dinv is None:
        Dinv = get_block_diag(lvl.A, blocksize=blocksize, inv_flag=True)
    if withrho:
        omega = omega / rho_block_D_inv_A(lvl.A, Dinv)
    smoother = partial(relaxation.block_jacobi, iterations=iterations, omega=omega, Dinv=Dinv,
                       blocksize=blocksize)
    update_wrapper(smoother, relaxation.block_jacobi)
    return smoother


def setup_sor(lvl, omega=0.5, iterations=DEFAULT_NITER, sweep=DEFAULT_SWEEP):
    smoother = partial(relaxation.sor, iterations=iterations, omega=omega, sweep=sweep)
    update_wrapper(smoother, relaxation.sor)
    return smoother


def setup_gauss_seidel_indexed(lvl, indices=None, iterations=DEFAULT_NITER, sweep=DEFAULT_SWEEP,
                               coloring="auto"):
    """Multi-colour Gauss-Seidel: ``relaxation.gauss_seidel_indexed`` over rows sorted by colour
    (``order = argsort(colours, kind='stable')``).  If ``indices`` is not
    given a vertex colouring of A's graph is computed here.  ``coloring='auto'``: natural-order greedy on
    big levels (its lattice-like colour classes keep the x gathers local: smallest-last was measured 30 %
    slower on the 8.4 M-row level of the 256^3 hierarchy) and smallest-last below 2 M rows, where the
    number of dependent waves, not bandwidth, decides (60 -> 52 colours on the densest level)."""
    if indices is None:
        from ..graph import vertex_coloring
        if coloring == "auto":
            coloring = "greedy" if lvl.A.shape[0] > 2_000_000 else "smallest_last"
        cache = getattr(lvl.A, "_b200_color_order", None)       # pre and post usually ask for the same list
        if cache is None or cache[0] != coloring:
            colors = vertex_coloring(lvl.A, method=coloring)
            cache = (coloring, np.argsort(colors, kind="stable").astype(np.int32))
            try:
                lvl.A._b200_color_order = cache
            except AttributeError:
                pass
        indices = cache[1]
    smoother = partial(relaxation.gauss_seidel_indexed, indices=np.asarray(indices, dtype=np.int32),
                       iterations=iterations, sweep=sweep)
    update_wrapper(smoother, relaxation.gauss_seidel_indexed)
    return smoother


def _polynomial_closure(name, coefficients, iterations):
    """The closure shape the reference stores for 'richardson' / 'chebyshev' (smoothing.py:611-618, :627-647):
    a plain function named after the registry key whose cell variables hold the polynomial coefficients."""
    coefficients = np.asarray(coefficients, dtype=np.float64)

    def smoother(A, x, b):
        relaxation.polynomial(A, x, b, coefficients=coefficients, iterations=iterations)
    smoother.__name__ = name
    smoother.__qualname__ = name
    return smoother


def setup_richardson(lvl, iterations=DEFAULT_NITER, omega=1.0):
    """x += omega / rho(A) * (b - A x)  (smoothing.py:611-618)."""
    omega = omega / approximate_spectral_radius(lvl.A)
    return _polynomial_closure("richardson", [omega], iterations)


def setup_chebyshev(lvl, lower_bound=1.0 / 30.0, upper_bound=1.1, degree=3, iterations=DEFAULT_NITER):
    """Chebyshev polynomial smoother on [lower_bound, upper_bound] * rho(A)  (smoothing.py:627-647)."""
    rho = approximate_spectral_radius(lvl.A)
    a = rho * lower_bound
    b = rho * upper_bound
    coefficients = -chebyshev_polynomial_coefficients(a, b, degree)[:-1]      # drop the constant coefficient
    return _polynomial_closure("chebyshev", coefficients, iterations)


def _extract_splitting(lvl):
    """F- and C-point lists from ``lvl.splitting`` (smoothing.py:55-72)."""
    try:
        splitting = lvl.splitting
    except AttributeError as exc:
        raise AttributeError("CF splitting is required in hierarchy.") from exc
    if splitting.dtype != bool:
        raise ValueError("CF splitting is required to be boolean.")
    Fpts = np.where(np.logical_not(splitting))[0].astype(dtype=int)
    Cpts = np.where(splitting)[0].astype(dtype=int)
    return Fpts, Cpts


def _setup_cf(fn, lvl, f_iterations, c_iterations, iterations, omega, withrho):
    if withrho:
        omega = omega / rho_D_inv_A(lvl.A)
    Fpts, Cpts = _extract_splitting(lvl)
    smoother = partial(fn, Cpts=Cpts, Fpts=Fpts, f_iterations=f_iterations, c_iterations=c_iterations,
                       iterations=iterations, omega=omega)
    update_wrapper(smoother, fn)
    return smoother


def setup_cf_jacobi(lvl, f_iterations=DEFAULT_NITER, c_iterations=DEFAULT_NITER, iterations=DEFAULT_NITER,
                    omega=1.0, withrho=False):
    """C-point Jacobi sweeps followed by F-point sweeps (smoothing.py:678-691)."""
    return _setup_cf(relaxation.cf_jacobi, lvl, f_iterations, c_iterations, iterations, omega, withrho)


def setup_fc_jacobi(lvl, f_iterations=DEFAULT_NITER, c_iterations=DEFAULT_NITER, iterations=DEFAULT_NITER,
                    omega=1.0, withrho=False):
    """F-point Jacobi sweeps followed by C-point sweeps (smoothing.py:694-707) -- AIR's post-smoother."""
    return _setup_cf(relaxation.fc_jacobi, lvl, f_iterations, c_iterations, iterations, omega, withrho)


def _setup_cf_block(fn_block_name, setup_point, lvl, f_iterations, c_iterations, iterations, omega, Dinv,
                    blocksize, withrho):
    if blocksize is None and Dinv is None:
        if sparse.issparse(lvl.A) and lvl.A.format == "csr":
            blocksize = 1
        elif sparse.issparse(lvl.A) and lvl.A.format == "bsr":
            blocksize = lvl.A.blocksize[0]
    elif blocksize is None:
        blocksize = Dinv.blocksize[1] if (sparse.issparse(Dinv) and Dinv.format == "bsr") else 1
    if (lvl.A.shape[0] % blocksize) != 0:
        raise ValueError("Blocksize does not divide size of matrix.")
    if len(lvl.splitting) * blocksize != lvl.A.shape[0]:
        raise ValueError("Blocksize not compatible with CF-splitting and matrix size.")
    if blocksize != 1:
        Fpts, Cpts = _extract_splitting(lvl)
        if Dinv is None:
            Dinv = get_block_diag(lvl.A, blocksize=blocksize, inv_flag=True)
        if withrho:
            omega = omega / rho_block_D_inv_A(lvl.A, Dinv)
        fn = getattr(relaxation, fn_block_name)
        smoother = partial(fn, Cpts=Cpts, Fpts=Fpts, f_iterations=f_iterations, c_iterations=c_iterations,
                           iterations=iterations, omega=omega, Dinv=Dinv, blocksize=blocksize)     # smoothing.py:741-751
        update_wrapper(smoother, fn)
        return smoother
    # blocksize 1: block Jacobi is point Jacobi; the reference forwards only iterations / omega / withrho
    # (smoothing.py:735-739), the registry name stays the block one
    smoother = setup_point(lvl, iterations=iterations, omega=omega, withrho=withrho)
    smoother.__name__ = fn_block_name
    return smoother


def setup_cf_block_jacobi(lvl, f_iterations=DEFAULT_NITER, c_iterations=DEFAULT_NITER, iterations=DEFAULT_NITER,
                          omega=1.0, Dinv=None, blocksize=None, withrho=False):
    """smoothing.py:710-751."""
    return _setup_cf_block("cf_block_jacobi", setup_cf_jacobi, lvl, f_iterations, c_iterations, iterations,
                           omega, Dinv, blocksize, withrho)


def setup_fc_block_jacobi(lvl, f_iterations=DEFAULT_NITER, c_iterations=DEFAULT_NITER, iterations=DEFAULT_NITER,
                          omega=1.0, Dinv=None, blocksize=None, withrho=False):
    """smoothing.py:754-791."""
    return _setup_cf_block("fc_block_jacobi", setup_fc_jacobi, lvl, f_iterations, c_iterations, iterations,
                           omega, Dinv, blocksize, withrho)


def setup_block_gauss_seidel(lvl, iterations=DEFAULT_NITER, sweep=DEFAULT_SWEEP, Dinv=None, blocksize=None):
    """Block Gauss-Seidel, the reference's default SA smoother (smoothing.py:582-608); blocksize 1 is plain
    Gauss-Seidel under the block name."""
    if blocksize is None and Dinv is None:
        if sparse.issparse(lvl.A) and lvl.A.format == "csr":
            blocksize = 1
        elif sparse.issparse(lvl.A) and lvl.A.format == "bsr":
            blocksize = lvl.A.blocksize[0]
    elif blocksize is None:
        blocksize = Dinv.shape[1]
    if blocksize == 1:
        smoother = setup_gauss_seidel(lvl, iterations=iterations, sweep=sweep)
        update_wrapper(smoother, relaxation.block_gauss_seidel)
        return smoother
    if Dinv is None:
        Dinv = get_block_diag(lvl.A, blocksize=blocksize, inv_flag=True)
    smoother = partial(relaxation.block_gauss_seidel, iterations=iterations, Dinv=Dinv, blocksize=blocksize,
                       sweep=sweep)
    update_wrapper(smoother, relaxation.block_gauss_seidel)
    return smoother


def _ne_closure(name, lvl, **params):
    """The closure shape the reference stores for the normal-equation smoothers (smoothing.py:641-675): a plain function
    named after the registry key; cell variables hold the level and the scalar parameters."""
    fn = getattr(relaxation, name)

    def smoother(A, x, b):
        fn(lvl.A, x, b, **params)
    smoother.__name__ = name
    smoother.__qualname__ = name
    smoother._ne_parameters = dict(params)
    return smoother


def setup_jacobi_ne(lvl, iterations=DEFAULT_NITER, omega=1.0, withrho=True):
    """Jacobi on the normal equations A A^H y = b (smoothing.py:641-651)."""
    if withrho:
        omega = omega / rho_D_inv_A(lvl.A) ** 2
    return _ne_closure("jacobi_ne", lvl, iterations=iterations, omega=omega)


def setup_gauss_seidel_ne(lvl, iterations=DEFAULT_NITER, sweep=DEFAULT_SWEEP, omega=1.0):
    """Gauss-Seidel on A A^H y = b (Kaczmarz; smoothing.py:654-663)."""
    return _ne_closure("gauss_seidel_ne", lvl, iterations=iterations, sweep=sweep, omega=omega)


def setup_gauss_seidel_nr(lvl, iterations=DEFAULT_NITER, sweep=DEFAULT_SWEEP, omega=1.0):
    """Gauss-Seidel on A^H A x = A^H b (smoothing.py:666-675)."""
    return _ne_closure("gauss_seidel_nr", lvl, iterations=iterations, sweep=sweep, omega=omega)


def normal_equation_closure_parameters(sm):
    """(name, {iterations, sweep, omega}) of a jacobi_ne / gauss_seidel_ne / gauss_seidel_nr smoother closure -- the
    reference's (parameters in cell variables, smoothing.py:641-675) or the ones built above -- else None."""
    name = getattr(sm, "__name__", None)
    if name not in ("jacobi_ne", "gauss_seidel_ne", "gauss_seidel_nr"):
        return None
    own = getattr(sm, "_ne_parameters", None)
    if own is not None:
        return name, dict(own)
    code, cells = getattr(sm, "__code__", None), getattr(sm, "__closure__", None)
    if code is None or cells is None:
        return None
    cv = {k: c.cell_contents for k, c in zip(code.co_freevars, cells)}
    return name, {k: cv[k] for k in ("iterations", "sweep", "omega") if k in cv}


def _level_csr(lvl):
    """The level operator as sorted CSR, kept on the level as ``Acsr`` (matrix_asformat, smoothing.py:441-491)."""
    if not hasattr(lvl, "Acsr"):
        lvl.Acsr = lvl.A if lvl.A.format == "csr" else lvl.A.tocsr()
    lvl.Acsr.sort_indices()
    return lvl.Acsr


def setup_schwarz(lvl, iterations=DEFAULT_NITER, subdomain=None, subdomain_ptr=None, inv_subblock=None,
                  inv_subblock_ptr=None, sweep=DEFAULT_SWEEP):
    """Overlapping multiplicative Schwarz (smoothing.py:509-526): subdomains and block inverses once, at setup."""
    A = _level_csr(lvl)
    params = dict(zip(("subdomain", "subdomain_ptr", "inv_subblock", "inv_subblock_ptr"),
                      relaxation.schwarz_parameters(A, subdomain, subdomain_ptr, inv_subblock, inv_subblock_ptr)))
    params.update(iterations=iterations, sweep=sweep)

    def smoother(A_, x, b):
        relaxation.schwarz(lvl.Acsr, x, b, **params)
    smoother.__name__ = smoother.__qualname__ = "schwarz"
    smoother._schwarz_parameters = params
    return smoother


def setup_strength_based_schwarz(lvl, iterations=DEFAULT_NITER, sweep=DEFAULT_SWEEP):
    """Schwarz whose subdomains are the rows of the strength matrix C (smoothing.py:529-548).  The reference
    rebuilds the block inverses on every application; they depend on A and C only, so they are built once here."""
    C = (lvl.C if hasattr(lvl, "C") else lvl.A).tocsr()
    C.sort_indices()
    smoother = setup_schwarz(lvl, iterations=iterations, subdomain=C.indices.copy(), subdomain_ptr=C.indptr.copy(),
                             sweep=sweep)
    # the registry key of THIS smoother (rebuild_smoother / hierarchy_io look it up by name and must re-derive the
    # subdomains from lvl.C, not fall back to A's sparsity pattern)
    smoother.__name__ = smoother.__qualname__ = "strength_based_schwarz"
    return smoother


def schwarz_closure_parameters(sm, A):
    """The parameters of a schwarz / strength_based_schwarz smoother closure -- the reference's (cell variables,
    smoothing.py:509-548) or the ones built above -- as a dict, else None.  A is the level operator (the
    strength-based closure of the reference carries the subdomains only; the inverses are built from A here)."""
    name = getattr(sm, "__name__", None)
    if name not in ("schwarz", "strength_based_schwarz"):
        return None
    own = getattr(sm, "_schwarz_parameters", None)
    if own is not None:
        return dict(own)
    code, cells = getattr(sm, "__code__", None), getattr(sm, "__closure__", None)
    if code is None or cells is None:
        return None
    cv = {k: c.cell_contents for k, c in zip(code.co_freevars, cells)}
    if name == "strength_based_schwarz":
        Acsr = sparse.csr_array(A)
        Acsr.sort_indices()
        cv["subdomain"], cv["subdomain_ptr"], cv["inv_subblock"], cv["inv_subblock_ptr"] = \
            relaxation.schwarz_parameters(Acsr, cv["subdomain"], cv["subdomain_ptr"])
    keys = ("iterations", "subdomain", "subdomain_ptr", "inv_subblock", "inv_subblock_ptr", "sweep")
    if any(k not in cv for k in keys):
        return None
    return {k: cv[k] for k in keys}


def setup_none(lvl):
    def none(A, x, b):
        pass
    return none


_REGISTER = {
    "gauss_seidel": setup_gauss_seidel,
    "jacobi": setup_jacobi,
    "block_jacobi": setup_block_jacobi,
    "sor": setup_sor,
    "gauss_seidel_indexed": setup_gauss_seidel_indexed,
    "multicolor_gauss_seidel": setup_gauss_seidel_indexed,
    "richardson": setup_richardson,
    "chebyshev": setup_chebyshev,
    "cf_jacobi": setup_cf_jacobi,
    "fc_jacobi": setup_fc_jacobi,
    "cf_block_jacobi": setup_cf_block_jacobi,
    "fc_block_jacobi": setup_fc_block_jacobi,
    "block_gauss_seidel": setup_block_gauss_seidel,
    "jacobi_ne": setup_jacobi_ne,
    "gauss_seidel_ne": setup_gauss_seidel_ne,
    "gauss_seidel_nr": setup_gauss_seidel_nr,
    "schwarz": setup_schwarz,
    "strength_based_schwarz": setup_strength_based_schwarz,
    "none": setup_none,
}

# in the reference's registry (smoothing.py:840-878) but outside the accelerated path
_OUT_OF_SCOPE = ["gmres", "cg", "cgne", "cgnr"]


def _setup_call(fn):
    if fn is None:
        fn = "none"
    if not isinstance(fn, str):
        raise ValueError(f"Input function must be a string or None: fn={fn}")
    if fn in _OUT_OF_SCOPE:
        raise NotImplementedError(f"smoother '{fn}' is not on the GPU hot path (no CPU fallback)")
    if fn not in _REGISTER:
        raise ValueError(f"Function {fn} does not have a setup")
    return _REGISTER[fn]


def _is_symmetric_pair(fn1, kw1, fn2, kw2):
    """Truth table of smoothing.py:226-262."""
    if kw1.get("iterations", DEFAULT_NITER) != kw2.get("iterations", DEFAULT_NITER):
        return False
    if (fn1, fn2) in (("cf_jacobi", "fc_jacobi"), ("fc_jacobi", "cf_jacobi"),
                      ("cf_block_jacobi", "fc_block_jacobi"), ("fc_block_jacobi", "cf_block_jacobi")):
        return (kw1.get("f_iterations", DEFAULT_NITER) == kw2.get("f_iterations", DEFAULT_NITER)
                and kw1.get("c_iterations", DEFAULT_NITER) == kw2.get("c_iterations", DEFAULT_NITER))
    if fn1 != fn2:
        return False
    if fn1 not in SYMMETRIC_RELAXATION:
        if fn1.startswith(("cf_", "fc_")):
            return False
        s1 = kw1.get("sweep", DEFAULT_SWEEP)
        s2 = kw2.get("sweep", DEFAULT_SWEEP)
        if (s1, s2) not in [("forward", "backward"), ("backward", "forward"), ("symmetric", "symmetric")]:
            return False
    return True


def change_smoothers(ml, presmoother, postsmoother):
    """Install pre/post smoothers on every level but the coarsest (smoothing.py:75-369)."""
    ml.symmetric_smoothing = True
    if isinstance(presmoother, (str, tuple)) or presmoother is None:
        presmoother = [presmoother]
    elif not isinstance(presmoother, list):
        raise ValueError('Unrecognized presmoother -- use a string:\n '
                         '"method" or ("method", opts) or list thereof.')
    if isinstance(postsmoother, (str, tuple)) or postsmoother is None:
        postsmoother = [postsmoother]
    elif not isinstance(postsmoother, list):
        raise ValueError('Unrecognized postsmoother -- use a string:\n '
                         '"method" or ("method", opts) or list thereof.')
    nlv = len(ml.levels) - 1
    for i in range(nlv):
        fn1, kw1 = _unpack_arg(presmoother[min(i, len(presmoother) - 1)])
        fn2, kw2 = _unpack_arg(postsmoother[min(i, len(postsmoother) - 1)])
        ml.levels[i].presmoother = _setup_call(fn1)(ml.levels[i], **kw1)
        ml.levels[i].postsmoother = _setup_call(fn2)(ml.levels[i], **kw2)
        if i < max(len(presmoother), len(postsmoother)) and not _is_symmetric_pair(fn1, kw1, fn2, kw2):
            ml.symmetric_smoothing = False
    if hasattr(ml, "_invalidate"):
        ml._invalidate()


def rebuild_smoother(lvl):
    """Rebuild the pre/post smoothers of a level from their registry names (smoothing.py:881-908)."""
    try:
        fn1 = lvl.presmoother.__name__
        fn2 = lvl.postsmoother.__name__
    except AttributeError as exc:
        raise AttributeError("The pre/post smoothers need to be functions.") from exc
    lvl.presmoother = _setup_call(fn1)(lvl)
    lvl.postsmoother = _setup_call(fn2)(lvl)


# --------------------------------------------------------------------------------------------
# closure -> engine descriptor
# --------------------------------------------------------------------------------------------
def describe(sm, A, keep):
    """Parse a level's smoother object into the C ``amgb_smoother`` descriptor.

    Accepts the closures of the reference (pyamg/relaxation/smoothing.py, see the descriptor
    table in SURVEY.md) and the ones built above: a ``functools.partial`` whose ``.func.__name__``
    names the relaxation routine, the no-op ``none`` function, or None.  Anything else -- the
    closure-based smoothers (richardson, chebyshev, schwarz, Krylov, ...) -- is outside the
    accelerated path: NotImplementedError, never a CPU fallback.
    """
    S = E.Smoother()
    S.kind, S.iterations, S.sweep, S.blocksize, S.omega = E.SM_NONE, 1, 0, 1, 1.0
    S.indices, S.n_indices, S.Dinv = None, 0, None
    if sm is None:
        return S
    S.indices2, S.n_indices2, S.f_iterations, S.c_iterations = None, 0, 1, 1
    S.coefficients, S.n_coefficients, S.reserved_ = None, 0, 0
    func = getattr(sm, "func", None)
    if func is None:
        if getattr(sm, "__name__", None) == "none":
            return S
        poly = polynomial_closure_parameters(sm)
        if poly is not None:
            coef = np.ascontiguousarray(poly[0], dtype=np.float64).reshape(-1)
            if coef.size < 1:
                raise ValueError("polynomial smoother without coefficients")
            keep.append(coef)
            S.kind, S.iterations = E.SM_POLYNOMIAL, int(poly[1])
            S.coefficients, S.n_coefficients = E.f64p(coef), coef.size
            return S
        ne = normal_equation_closure_parameters(sm)
        if ne is not None:
            name, kw = ne
            if getattr(A, "format", "csr") == "bsr" and A.blocksize != (1, 1):
                raise NotImplementedError("normal-equation smoothers on a block operator are not on the GPU hot path")
            S.kind = {"jacobi_ne": E.SM_JACOBI_NE, "gauss_seidel_ne": E.SM_GAUSS_SEIDEL_NE,
                      "gauss_seidel_nr": E.SM_GAUSS_SEIDEL_NR}[name]
            S.iterations = int(kw.get("iterations", 1))
            S.omega = float(np.real(kw.get("omega", 1.0)))
            sweep = kw.get("sweep", "forward")
            if sweep not in E.SWEEPS:
                raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
            S.sweep = E.SWEEPS[sweep]
            Dinv = np.ascontiguousarray(get_diagonal(A, norm_eq=1 if name == "gauss_seidel_nr" else 2, inv=True))
            keep.append(Dinv)
            S.Dinv = E.f64p(Dinv)
            return S
        sz = schwarz_closure_parameters(sm, A)
        if sz is not None:
            if sz["sweep"] not in E.SWEEPS:
                raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
            relaxation._fill_schwarz(S, keep, sz["subdomain"], sz["subdomain_ptr"], sz["inv_subblock"],
                                     sz["iterations"], sz["sweep"])
            return S
        raise NotImplementedError(
            f"smoother {getattr(sm, '__name__', sm)!r} is a closure the GPU engine cannot introspect; "
            "supported: jacobi, gauss_seidel, gauss_seidel_indexed (multi-colour), block_jacobi, sor, richardson, "
            "chebyshev, cf_jacobi, fc_jacobi, jacobi_indexed, block_gauss_seidel, None")
    name = func.__name__
    kw = dict(sm.keywords)
    S.iterations = int(kw.get("iterations", 1))
    if name == "jacobi":
        S.kind = E.SM_JACOBI
        S.omega = float(np.real(kw.get("omega", 1.0)))
    elif name in ("gauss_seidel", "sor", "gauss_seidel_indexed"):
        S.kind = E.SM_GAUSS_SEIDEL
        sweep = kw.get("sweep", "forward")
        if sweep not in E.SWEEPS:
            raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
        S.sweep = E.SWEEPS[sweep]
        S.omega = float(np.real(kw.get("omega", 1.0))) if name != "gauss_seidel_indexed" else 1.0
        if getattr(A, "format", None) == "bsr":
            # reference quirk: on a BSR operator gauss_seidel / sor call bsr_gauss_seidel, which has no omega
            # (relaxation.py:343-346) -- every SA coarse level is BSR, so 'sor' there is plain Gauss-Seidel
            S.omega = 1.0
        if name == "gauss_seidel_indexed":
            idx = np.ascontiguousarray(np.asarray(kw["indices"], dtype="intc"), dtype=np.int32)
            keep.append(idx)
            S.indices = E.i32p(idx)
            S.n_indices = len(idx)
    elif name == "block_jacobi":
        S.kind = E.SM_BLOCK_JACOBI
        S.omega = float(np.real(kw.get("omega", 1.0)))
        bs = int(kw.get("blocksize", 1))
        Dinv = kw.get("Dinv", None)
        if Dinv is None:
            Dinv = get_block_diag(A, blocksize=bs, inv_flag=True)
        Dinv = np.ascontiguousarray(Dinv, dtype=np.float64)
        if Dinv.shape != (A.shape[0] // bs, bs, bs):
            raise ValueError("Dinv and A have incompatible dimensions")
        keep.append(Dinv)
        S.blocksize = bs
        S.Dinv = E.f64p(Dinv.reshape(-1))
    elif name == "polynomial":
        coef = np.ascontiguousarray(np.real(np.asarray(kw["coefficients"])), dtype=np.float64).reshape(-1)
        if coef.size < 1:
            raise ValueError("polynomial smoother without coefficients")
        keep.append(coef)
        S.kind = E.SM_POLYNOMIAL
        S.coefficients, S.n_coefficients = E.f64p(coef), coef.size
    elif name in ("jacobi_indexed", "cf_jacobi", "fc_jacobi"):
        S.omega = float(np.real(kw.get("omega", 1.0)))
        if name == "jacobi_indexed":
            S.kind = E.SM_JACOBI_INDEXED
            idx = np.ascontiguousarray(np.asarray(kw["indices"]), dtype=np.int32)
        else:
            S.kind = E.SM_CF_JACOBI if name == "cf_jacobi" else E.SM_FC_JACOBI
            idx = np.ascontiguousarray(np.asarray(kw["Cpts"]), dtype=np.int32)
            idx2 = np.ascontiguousarray(np.asarray(kw["Fpts"]), dtype=np.int32)
            keep.append(idx2)
            S.indices2, S.n_indices2 = E.i32p(idx2), len(idx2)
            S.f_iterations = int(kw.get("f_iterations", 1))
            S.c_iterations = int(kw.get("c_iterations", 1))
            if getattr(A, "format", "csr") == "bsr" and A.blocksize != (1, 1):
                raise NotImplementedError("CF Jacobi on a BSR operator (bsr_jacobi_indexed) is not on the GPU hot path")
        keep.append(idx)
        S.indices, S.n_indices = E.i32p(idx), len(idx)
    elif name in ("cf_block_jacobi", "fc_block_jacobi"):
        S.kind = E.SM_CF_BLOCK_JACOBI if name == "cf_block_jacobi" else E.SM_FC_BLOCK_JACOBI
        S.omega = float(np.real(kw.get("omega", 1.0)))
        bs = int(kw.get("blocksize", 1))
        Dinv = kw.get("Dinv", None)
        if Dinv is None:
            Dinv = get_block_diag(A, blocksize=bs, inv_flag=True)
        Dinv = np.ascontiguousarray(Dinv, dtype=np.float64)
        if Dinv.shape != (A.shape[0] // bs, bs, bs):
            raise ValueError("Dinv and A have incompatible dimensions")
        idx = np.ascontiguousarray(np.asarray(kw["Cpts"]), dtype=np.int32)
        idx2 = np.ascontiguousarray(np.asarray(kw["Fpts"]), dtype=np.int32)
        keep += [Dinv, idx, idx2]
        S.blocksize, S.Dinv = bs, E.f64p(Dinv.reshape(-1))
        S.indices, S.n_indices = E.i32p(idx), len(idx)
        S.indices2, S.n_indices2 = E.i32p(idx2), len(idx2)
        S.f_iterations, S.c_iterations = int(kw.get("f_iterations", 1)), int(kw.get("c_iterations", 1))
    elif name == "block_gauss_seidel":
        S.kind = E.SM_BLOCK_GAUSS_SEIDEL
        sweep = kw.get("sweep", "forward")
        if sweep not in E.SWEEPS:
            raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
        S.sweep = E.SWEEPS[sweep]
        bs = int(kw.get("blocksize", 1))
        Dinv = kw.get("Dinv", None)
        if Dinv is None:
            Dinv = get_block_diag(A, blocksize=bs, inv_flag=True)
        Dinv = np.ascontiguousarray(Dinv, dtype=np.float64)
        if Dinv.shape[0] != A.shape[0] // bs:
            raise ValueError("Dinv and A have incompatible dimensions")
        if Dinv.shape[1] != bs or Dinv.shape[2] != bs:
            raise ValueError("Dinv and blocksize are incompatible")
        keep.append(Dinv)
        S.blocksize = bs
        S.Dinv = E.f64p(Dinv.reshape(-1))
    else:
        raise NotImplementedError(f"smoother '{name}' is not on the GPU hot path (no CPU fallback)")
    return S


def polynomial_closure_parameters(sm):
    """(coefficients, iterations) of a 'richardson' / 'chebyshev' smoother closure, or None.

    The reference keeps these smoothers' parameters in closure cells rather than ``partial`` keywords
    (smoothing.py:611-618: ``omega``, ``iterations``; :627-647: ``coefficients``, ``iterations``); the cells are
    read here, which is how a hierarchy built by the reference is adopted without re-running its setup."""
    name = getattr(sm, "__name__", None)
    code = getattr(sm, "__code__", None)
    cells = getattr(sm, "__closure__", None)
    if name not in ("richardson", "chebyshev") or code is None or cells is None:
        return None
    cv = {k: c.cell_contents for k, c in zip(code.co_freevars, cells)}
    if "coefficients" in cv:
        coef = cv["coefficients"]
    elif "omega" in cv:
        coef = [cv["omega"]]
    else:
        return None
    return np.real(np.asarray(coef, dtype=np.float64)), int(cv.get("iterations", 1))
