"""CPU tests of the host side: C-ABI export list, smoother-closure parsing, the smoother factory's
truth tables, hierarchy (de)serialisation, reporting methods, and loud failure without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import pyamg_b200
from pyamg_b200 import _engine as E
from pyamg_b200.relaxation import smoothing, relaxation
from conftest import GOLDEN, GOLDEN_ALL, ROOT
from kats import poisson1d


def test_c_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pyamg_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(amgb_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    L = E.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(E.SYMBOLS) == declared
    assert L.amgb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    if E.lib().amgb_device_count() > 0:
        pytest.skip("GPU present")
    A = poisson1d(4)
    with pytest.raises(E.EngineError):
        relaxation.jacobi(A, np.zeros(4), np.ones(4))
    lvl = pyamg_b200.MultilevelSolver.Level()
    lvl.A = A
    with pytest.raises(E.EngineError):
        pyamg_b200.MultilevelSolver([lvl]).solve(np.ones(4))
    h = ctypes.c_void_p()
    assert E.lib().amgb_hierarchy_create(0, ctypes.byref(h)) != 0
    assert E.lib().amgb_last_error()


def test_validation_precedes_device_use():
    """make_system's contract (relaxation.py:15-97) is enforced on the host, GPU or not."""
    A = poisson1d(4)
    with pytest.raises(ValueError):
        relaxation.jacobi(A, [0.0] * 4, np.zeros(4))
    with pytest.raises(TypeError):
        relaxation.jacobi(A, np.zeros(4, dtype=np.float32), np.zeros(4))
    with pytest.raises(ValueError):
        relaxation.gauss_seidel(A, np.zeros(8)[::2], np.zeros(4))
    with pytest.raises(ValueError):
        relaxation.gauss_seidel(A, np.zeros(4), np.zeros(4), sweep="sideways")
    with pytest.raises(ValueError):
        relaxation.block_jacobi(A, np.zeros(4), np.zeros(4), Dinv=np.zeros((3, 2, 2)), blocksize=2)


@pytest.mark.parametrize("name", GOLDEN)
def test_closure_parsing_of_reference_built_hierarchies(name, load_golden):
    ml, _ = load_golden(name)
    for lvl in ml.levels[:-1]:
        for which in ("presmoother", "postsmoother"):
            keep = []
            S = smoothing.describe(getattr(lvl, which), lvl.A, keep)
            sm = getattr(lvl, which)
            if getattr(sm, "func", None) is None:
                assert S.kind == E.SM_NONE
                continue
            fn = sm.func.__name__
            assert S.iterations == sm.keywords.get("iterations", 1)
            if fn == "jacobi":
                assert S.kind == E.SM_JACOBI and S.omega == pytest.approx(float(sm.keywords["omega"]))
            elif fn == "block_jacobi":
                assert S.kind == E.SM_BLOCK_JACOBI and S.blocksize == sm.keywords["blocksize"]
            else:
                assert S.kind == E.SM_GAUSS_SEIDEL
                assert S.sweep == E.SWEEPS[sm.keywords.get("sweep", "forward")]
                assert (S.n_indices > 0) == ("indices" in sm.keywords)


def test_unsupported_smoothers_fail_loudly():
    def richardson(A, x, b):
        pass
    with pytest.raises(NotImplementedError):
        smoothing.describe(richardson, poisson1d(3), [])
    with pytest.raises(NotImplementedError):
        smoothing._setup_call("cgnr")
    with pytest.raises(ValueError):
        smoothing._setup_call("no_such_smoother")
    with pytest.raises(NotImplementedError):
        pyamg_b200.coarse_grid_solver("cg")


def _two_level():
    A = poisson1d(8)
    P = sp.csr_array(np.kron(np.eye(4), np.ones((2, 1))))
    l0, l1 = pyamg_b200.MultilevelSolver.Level(), pyamg_b200.MultilevelSolver.Level()
    l0.A, l0.P = A, P
    l1.A = sp.csr_array(P.T @ A @ P)
    return pyamg_b200.MultilevelSolver([l0, l1])


def test_change_smoothers_grammar_and_symmetry_flag():
    """pyamg/relaxation/tests/test_smoothing.py:94-145 (truth table, __name__ kept)."""
    ml = _two_level()
    assert hasattr(ml.levels[0], "R")                       # multilevel.py:180-182
    cases = [(("gauss_seidel", {"sweep": "symmetric"}), ("gauss_seidel", {"sweep": "symmetric"}), True),
             (("gauss_seidel", {"sweep": "forward"}), ("gauss_seidel", {"sweep": "backward"}), True),
             (("gauss_seidel", {"sweep": "forward"}), ("gauss_seidel", {"sweep": "forward"}), False),
             ("jacobi", "jacobi", True),
             (("jacobi", {"iterations": 2}), "jacobi", False),
             ("jacobi", "gauss_seidel", False),
             (None, None, True)]
    for pre, post, sym in cases:
        pyamg_b200.change_smoothers(ml, pre, post)
        assert ml.symmetric_smoothing is sym, (pre, post)
    pyamg_b200.change_smoothers(ml, "block_jacobi", ("jacobi", {"omega": 4.0 / 3.0}))
    assert ml.levels[0].presmoother.__name__ == "block_jacobi"      # rewrapped Jacobi keeps the registry name
    assert ml.levels[0].presmoother.func is relaxation.jacobi
    w = ml.levels[0].postsmoother.keywords["omega"]
    rho = smoothing.rho_D_inv_A(ml.levels[0].A)
    assert w == pytest.approx(4.0 / 3.0 / rho) and 1.7 < rho < 2.0   # rho(D^-1 A) of 1-D Poisson -> 2
    with pytest.raises(ValueError):
        pyamg_b200.change_smoothers(ml, 3.14, None)
    smoothing.rebuild_smoother(ml.levels[0])
    assert ml.levels[0].presmoother.__name__ == "block_jacobi"


@pytest.mark.parametrize("name", GOLDEN_ALL)
def test_hierarchy_io_roundtrip_and_reports(name, load_golden, tmp_path):
    from pyamg_b200.hierarchy_io import save_hierarchy, load_hierarchy
    ml, ex = load_golden(name)
    p = str(tmp_path / "h.npz")
    save_hierarchy(p, ml, extra={"b": ex["b"]})
    ml2, ex2 = load_hierarchy(p)
    assert np.array_equal(ex2["b"], ex["b"])
    assert len(ml2.levels) == len(ml.levels)
    for a, b in zip(ml.levels, ml2.levels):
        assert (a.A != b.A).nnz == 0 and a.A.format == b.A.format
        if hasattr(a, "P"):
            assert (a.P != b.P).nnz == 0 and (a.R != b.R).nnz == 0
            assert a.presmoother.__name__ == b.presmoother.__name__
    assert np.array_equal(ml.coarse_solver.P, ml2.coarse_solver.P)
    text = repr(ml)
    assert "MultilevelSolver" in text and f"Number of Levels:     {len(ml.levels)}" in text
    nnz = [lv.A.nnz for lv in ml.levels]
    assert ml.operator_complexity() == pytest.approx(sum(nnz) / nnz[0])
    assert ml.grid_complexity() == pytest.approx(sum(lv.A.shape[0] for lv in ml.levels) / ml.levels[0].A.shape[0])
    assert ml.cycle_complexity("W") >= ml.cycle_complexity("F") >= ml.cycle_complexity("V")
    with pytest.raises(TypeError):
        ml.cycle_complexity("Z")


@pytest.mark.parametrize("seed,symmetric,kind", [(1, True, "perm"), (2, False, "perm"), (3, False, "repeats"),
                                                 (4, True, "natural"), (5, False, "subset")])
def test_wave_schedule_reproduces_the_sequential_sweep(seed, symmetric, kind):
    """The engine's dependency-wave rule (csrc/engine.cu build_waves == amgb_wave_schedule): executing the
    waves in order -- rows of one wave in ANY order, here reversed -- gives exactly the sequential
    gauss_seidel_indexed sweep, for symmetric and non-symmetric patterns, permutations, repeated and partial
    row lists; and no wave contains two rows that reference each other."""
    import oracle
    from pyamg_b200.dist import wave_schedule
    rng = np.random.default_rng(seed)
    n = 120
    M = sp.random(n, n, density=0.06, random_state=seed, format="csr")
    if symmetric:
        M = (M + M.T).tocsr()
    A = (M + sp.diags_array(4.0 + rng.random(n))).tocsr()
    A.indptr, A.indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    if kind == "perm":
        lst = rng.permutation(n).astype(np.int32)
    elif kind == "repeats":
        lst = rng.integers(0, n, size=2 * n).astype(np.int32)
    elif kind == "subset":
        lst = rng.permutation(n)[: n // 3].astype(np.int32)
    else:
        lst = None
    wave, nw = wave_schedule(A, lst)
    rows = np.arange(n, dtype=np.int32) if lst is None else lst
    assert wave.min() == 1 and wave.max() == nw
    x0, b = rng.random(n), rng.random(n)
    xs = x0.copy()
    oracle.gauss_seidel_indexed(A, xs, b, rows, sweep="forward")
    xw = x0.copy()
    pattern = (A != 0).astype(np.int8)
    for w in range(1, nw + 1):
        members = rows[wave == w][::-1]
        sub = pattern[members][:, members]
        assert sub.nnz == len(np.unique(members)) or len(np.unique(members)) < len(members) or \
            (sub - sp.diags_array(sub.diagonal())).nnz == 0
        oracle.gauss_seidel_indexed(A, xw, b, members.astype(np.int32), sweep="forward")
    assert np.array_equal(xw, xs)
    # the backward sweep is the waves in reverse order
    xs = x0.copy()
    oracle.gauss_seidel_indexed(A, xs, b, rows, sweep="backward")
    xw = x0.copy()
    for w in range(nw, 0, -1):
        oracle.gauss_seidel_indexed(A, xw, b, rows[wave == w].astype(np.int32), sweep="forward")
    assert np.array_equal(xw, xs)


@pytest.mark.parametrize("G,T,RMAX", [(1, 224, 64), (2, 224, 64), (4, 256, 64), (1, 512, 128), (32, 224, 64)])
def test_tile_builder_invariants(G, T, RMAX):
    """The TMA tile list (csrc/engine.cu build_tiles): tiles are consecutive whole-row ranges covering every
    row once, hold <= T entries and <= RMAX rows unless they are a single over-long row, never cross a break
    (Gauss-Seidel wave boundary), empty break ranges get empty tile ranges, and multi-pass tiles are cut to a
    multiple of the rows reduced per pass (32/G)."""
    rng = np.random.default_rng(G * 1000 + T)
    n = 3000
    lens = rng.integers(0, 40, size=n)
    lens[rng.integers(0, n, size=5)] = T + 50            # rows longer than a tile
    lens[100:140] = 0                                    # a run of empty rows
    Ap = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cuts = np.unique(np.concatenate([[0, n], rng.integers(0, n, size=12)]))
    breaks = np.sort(np.concatenate([cuts, cuts[3:5]])).astype(np.int64)       # duplicated cuts = empty waves
    for use_breaks in (False, True):
        cap = n + 8
        row0, nz0 = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32)
        tp = np.empty(len(breaks), dtype=np.int32)
        nt = ctypes.c_int32(0)
        E.check(E.lib().amgb_debug_build_tiles(n, E.i32p(Ap), G, breaks.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
                                               if use_breaks else None, len(breaks) - 1, T, RMAX, E.i32p(row0),
                                               E.i32p(nz0), cap, E.i32p(tp) if use_breaks else None, ctypes.byref(nt)))
        nt = nt.value
        r, z = row0[:nt + 1], nz0[:nt + 1]
        assert r[0] == 0 and r[-1] == n and np.all(np.diff(r) >= 1)
        assert np.array_equal(z, Ap[r])
        rows, ents = np.diff(r), np.diff(z)
        assert np.all((ents <= T) | (rows == 1))
        assert np.all(rows <= RMAX)
        if use_breaks:
            assert np.all(np.isin(breaks, r))                         # every break is a tile start
            for w in range(len(breaks) - 1):
                assert r[tp[w]] == breaks[w] or breaks[w] == breaks[w + 1]
                covered = r[tp[w + 1]] - r[tp[w]]
                assert covered == breaks[w + 1] - breaks[w]
        rpp = 32 // G
        inner = (rows > rpp) & (ents + lens[np.minimum(r[1:], n - 1)] <= T) & (rows < RMAX)
        if not use_breaks:
            assert np.all(rows[inner][:-1] % rpp == 0)


def test_from_pyamg_adopts_a_reference_style_solver_object(load_golden):
    """Drop-in boundary: `from_pyamg` takes any object with pyamg.MultilevelSolver's attributes -- here a
    stand-in whose levels/closures/coarse solver were produced by the real reference (golden) -- shares the
    Level objects, keeps symmetric_smoothing, parses the coarse-solver name and re-uses the cached pinv."""
    src, _ = load_golden("cfg5_sa_bjacobi_elasticity")

    class RefCoarse:                       # pyamg's GenericSolver: name() -> repr(solver), cached .P after first use
        P = src.coarse_solver.P

        @classmethod
        def name(cls):
            return "'pinv'"

    class RefSolver:
        levels = src.levels
        coarse_solver = RefCoarse()
        symmetric_smoothing = True

    ml = pyamg_b200.MultilevelSolver.from_pyamg(RefSolver())
    assert ml.levels[0] is src.levels[0] and len(ml.levels) == len(src.levels)
    assert ml.symmetric_smoothing is True
    assert ml.coarse_solver.name() == "'pinv'" and ml.coarse_solver.P is not None
    assert np.array_equal(ml.coarse_solver.dense_operator(ml.levels[-1].A), src.coarse_solver.P)
    assert "Number of Levels:     3" in repr(ml)

    class Unsupported(RefSolver):
        class coarse_solver:               # noqa: N801 - mimics an instance with a name() method
            @staticmethod
            def name():
                return "'cg'"
    with pytest.raises(NotImplementedError):
        pyamg_b200.MultilevelSolver.from_pyamg(Unsupported())
    # tuple form ('pinv', {...}) as the reference's coarse_grid_solver accepts it (multilevel.py:710-715)
    cs = pyamg_b200.coarse_grid_solver(("pinv", {"atol": 1e-12}))
    assert cs.name() == "('pinv', {'atol': 1e-12})"


def test_widened_smoother_descriptors_are_parsed(load_golden):
    """Closures of the reference's richardson / chebyshev (parameters in cell variables), cf/fc Jacobi partials and
    block Gauss-Seidel partials -> engine descriptors."""
    ml, _ = load_golden("cfg7_sa_cheby_richardson_poisson2d")
    keep = []
    S = smoothing.describe(ml.levels[0].presmoother, ml.levels[0].A, keep)
    assert S.kind == E.SM_POLYNOMIAL and S.n_coefficients == 3 and S.iterations == 1
    S = smoothing.describe(ml.levels[0].postsmoother, ml.levels[0].A, keep)
    assert S.kind == E.SM_POLYNOMIAL and S.n_coefficients == 1 and S.iterations == 2
    assert ml.levels[0].presmoother.__name__ == "chebyshev" and ml.levels[0].postsmoother.__name__ == "richardson"
    ml, _ = load_golden("cfg8_rs_cfjacobi_poisson3d")
    lvl = ml.levels[0]
    S = smoothing.describe(lvl.presmoother, lvl.A, keep)
    assert S.kind == E.SM_CF_JACOBI and S.f_iterations == 2 and S.c_iterations == 1 and S.omega == pytest.approx(0.8)
    assert S.n_indices + S.n_indices2 == lvl.A.shape[0]
    S = smoothing.describe(lvl.postsmoother, lvl.A, keep)
    assert S.kind == E.SM_FC_JACOBI and S.c_iterations == 2 and S.iterations == 2
    ml, _ = load_golden("cfg9_air_fcjacobi_advection2d")
    assert smoothing.describe(ml.levels[0].presmoother, ml.levels[0].A, keep).kind == E.SM_NONE
    assert smoothing.describe(ml.levels[0].postsmoother, ml.levels[0].A, keep).f_iterations == 2
    ml, _ = load_golden("cfg10_sa_bgs_elasticity")
    S = smoothing.describe(ml.levels[0].presmoother, ml.levels[0].A, keep)
    assert S.kind == E.SM_BLOCK_GAUSS_SEIDEL and S.blocksize == 2 and S.sweep == E.SWEEPS["symmetric"]
    assert smoothing.describe(ml.levels[1].presmoother, ml.levels[1].A, keep).blocksize == 3
    # still outside the accelerated path: loud, never a CPU fallback
    with pytest.raises(NotImplementedError):
        smoothing._setup_call("cgne")


def test_chebyshev_coefficients_reference_doctest():
    """pyamg/relaxation/chebyshev.py:29-31."""
    from pyamg_b200.relaxation.chebyshev import chebyshev_polynomial_coefficients
    c = chebyshev_polynomial_coefficients(1.0, 2.0, 3)
    assert np.allclose(c, [-0.32323232, 1.45454545, -2.12121212, 1.0], atol=5e-9)
    with pytest.raises(ValueError):
        chebyshev_polynomial_coefficients(2.0, 1.0, 3)


def test_coarse_solver_spec_roundtrip(load_golden, tmp_path):
    """Relaxation coarse solvers keep their keyword arguments through save / load (the reference's name() drops
    them: multilevel.py:820-823) and show up in the engine descriptor."""
    from pyamg_b200.hierarchy_io import save_hierarchy, load_hierarchy
    from pyamg_b200.multilevel import coarse_solver_spec
    ml, ex = load_golden("cfg11_rs_gs_coarse_relaxation")
    assert coarse_solver_spec(ml.coarse_solver) == ("gauss_seidel", {"iterations": 4, "sweep": "symmetric"})
    assert ml.coarse_solver.relaxation == ("gauss_seidel", {"iterations": 4, "sweep": "symmetric"})
    p = str(tmp_path / "h.npz")
    save_hierarchy(p, ml)
    ml2, _ = load_hierarchy(p)
    assert coarse_solver_spec(ml2.coarse_solver) == coarse_solver_spec(ml.coarse_solver)
    S = smoothing.describe(ml.coarse_solver.smoother(ml.levels[-1].A), ml.levels[-1].A, [])
    assert S.kind == E.SM_GAUSS_SEIDEL and S.iterations == 4 and S.sweep == E.SWEEPS["symmetric"]
    assert pyamg_b200.coarse_grid_solver("jacobi").relaxation == ("jacobi", {"iterations": 10})   # default: 10 sweeps
    assert coarse_solver_spec(pyamg_b200.coarse_grid_solver("pinv")) == "pinv"


def test_normal_equation_closures_are_parsed(load_golden):
    """The reference keeps jacobi_ne / gauss_seidel_ne / gauss_seidel_nr parameters in closure cells
    (smoothing.py:641-675): they survive adoption, save / load and reach the engine descriptor with Dinv = the inverse
    diagonal of A A^H (NE) or A^H A (NR)."""
    from pyamg_b200.util import get_diagonal
    ml, _ = load_golden("cfg13_rs_gsnr_gsne_advdiff2d")
    lvl = ml.levels[0]
    keep = []
    S = smoothing.describe(lvl.presmoother, lvl.A, keep)
    assert S.kind == E.SM_GAUSS_SEIDEL_NR and S.sweep == E.SWEEPS["symmetric"] and S.iterations == 1
    assert np.allclose(np.ctypeslib.as_array(S.Dinv, shape=(lvl.A.shape[0],)), 1.0 / (lvl.A.multiply(lvl.A)).sum(axis=0))
    S = smoothing.describe(lvl.postsmoother, lvl.A, keep)
    assert S.kind == E.SM_GAUSS_SEIDEL_NE and S.iterations == 2 and S.omega == pytest.approx(0.9)
    assert np.allclose(get_diagonal(lvl.A, norm_eq=2), np.asarray((lvl.A.multiply(lvl.A)).sum(axis=1)).ravel())
    ml, _ = load_golden("cfg14_sa_jacobine_poisson2d")
    S = smoothing.describe(ml.levels[0].postsmoother, ml.levels[0].A, keep)
    assert S.kind == E.SM_JACOBI_NE and S.iterations == 2 and 0 < S.omega < 4.0 / 3.0


def test_schwarz_closures_and_parameters(load_golden, tmp_path):
    """The reference keeps the Schwarz subdomains and block inverses in closure cells (smoothing.py:509-548): they
    reach the engine descriptor (indices = rows, indices2 = offsets, Dinv = inverses) and survive save / load;
    schwarz_parameters restates relaxation.py:1002-1078 incl. its cache-on-the-matrix rules."""
    import oracle
    from pyamg_b200.hierarchy_io import save_hierarchy, load_hierarchy
    from pyamg_b200.relaxation import relaxation
    ml, _ = load_golden("cfg15_sa_schwarz_aniso2d")
    lvl = ml.levels[0]
    keep = []
    S = smoothing.describe(lvl.presmoother, lvl.A, keep)
    assert S.kind == E.SM_SCHWARZ and S.sweep == E.SWEEPS["symmetric"] and S.iterations == 1
    assert S.n_indices2 == lvl.A.shape[0] + 1 and S.n_indices == lvl.A.nnz          # default: the sparsity patterns
    S2 = smoothing.describe(lvl.postsmoother, lvl.A, keep)
    assert S2.kind == E.SM_SCHWARZ and S2.sweep == E.SWEEPS["backward"] and S2.iterations == 2
    sb, _ = load_golden("cfg16_sa_strength_schwarz_aniso2d")
    S3 = smoothing.describe(sb.levels[0].postsmoother, sb.levels[0].A, keep)
    assert S3.kind == E.SM_SCHWARZ and S3.iterations == 2
    assert S3.n_indices2 == S.n_indices2 and S3.n_indices < S.n_indices             # strength pattern: fewer entries
    p = str(tmp_path / "h.npz")
    save_hierarchy(p, ml)
    ml2, _ = load_hierarchy(p)
    a, b = lvl.postsmoother._schwarz_parameters, ml2.levels[0].postsmoother._schwarz_parameters
    assert all(np.array_equal(a[k], b[k]) for k in ("subdomain", "subdomain_ptr", "inv_subblock", "inv_subblock_ptr"))
    # parameters: product == oracle == dense inverse of the blocks
    A = sp.csr_array(lvl.A).copy()
    A.sort_indices()
    mine = relaxation.schwarz_parameters(A)
    theirs = oracle.schwarz_parameters(sp.csr_array(lvl.A).copy())
    assert all(np.array_equal(u, v) for u, v in zip(mine, theirs))
    d = 17
    rows = mine[0][mine[1][d]:mine[1][d + 1]]
    blk = A.toarray()[np.ix_(rows, rows)]
    assert np.allclose(mine[2][mine[3][d]:mine[3][d + 1]].reshape(len(rows), -1) @ blk, np.eye(len(rows)), atol=1e-10)
    # cache rules: no subdomains named -> the cached set; other subdomains -> rebuilt and re-cached
    assert relaxation.schwarz_parameters(A) is mine
    sub, ptr = np.arange(A.shape[0], dtype=np.int32), np.arange(0, A.shape[0] + 1, 2, dtype=np.int32)
    other = relaxation.schwarz_parameters(A, sub, ptr)
    assert other is not mine and relaxation.schwarz_parameters(A) is other
    # singular block: pseudo-inverse (gelss with rcond = 1e6 eps), not an exception
    Z = sp.csr_array(np.array([[1.0, 1.0], [1.0, 1.0]]))
    T = relaxation.schwarz_parameters(Z, np.array([0, 1], dtype=np.int32), np.array([0, 2], dtype=np.int32))[2]
    assert np.allclose(T.reshape(2, 2), np.full((2, 2), 0.25))
