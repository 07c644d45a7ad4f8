"""GPU parity tests: the CUDA engine (through the C ABI of libpyamg_b200.so) against the oracle
and against the golden vectors the real reference produced.

Tolerance: ||x_gpu - x_ref|| / ||x_ref|| < 1e-12 in fp64 (BASELINE.json north_star); the only
difference to the reference is the summation order inside a row (warp-shuffle tree vs sequential).
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
import pyamg_b200
from pyamg_b200.relaxation import relaxation as gpu_relax
from conftest import GOLDEN, relerr
from kats import poisson1d, run_kats, gs_convergence_case, dense_gs_gold, block_jacobi_gold

pytestmark = pytest.mark.gpu
TOL = 1e-12


# ------------------------------------------------------------------ full cycles vs the reference
@pytest.mark.parametrize("name", GOLDEN)
def test_vcycle_matches_reference_golden(name, load_golden):
    ml, ex = load_golden(name)
    res = []
    x = ml.solve(ex["b"], tol=0, maxiter=len(ex["residuals"]) - 1, residuals=res)
    assert relerr(x, ex["x_ref"]) < TOL
    assert len(res) == len(ex["residuals"])
    assert np.allclose(res, ex["residuals"], rtol=1e-9, atol=1e-13 * ex["residuals"][0])
    assert ml.last_launches() > 0


@pytest.mark.parametrize("name", GOLDEN)
def test_w_and_f_cycles_match_reference_golden(name, load_golden):
    ml, ex = load_golden(name)
    assert relerr(ml.solve(ex["b"], tol=0, maxiter=2, cycle="W"), ex["x_ref_W"]) < TOL
    assert relerr(ml.solve(ex["b"], tol=0, maxiter=2, cycle="F"), ex["x_ref_F"]) < TOL


@pytest.mark.parametrize("name", GOLDEN)
def test_tolerance_stop_and_x0_match_reference_golden(name, load_golden):
    """x0 given, tol>0: same iteration count, same info flag, same iterate (multilevel.py:574-582)."""
    ml, ex = load_golden(name)
    res = []
    x0 = ex["x0"].copy()
    x, info = ml.solve(ex["b"], x0=x0, tol=1e-6, maxiter=50, residuals=res, return_info=True)
    assert np.array_equal(x0, ex["x0"])            # x0 is copied, not mutated (:467)
    assert info == int(ex["info_tol"][0])
    assert len(res) == len(ex["residuals_tol"])
    assert relerr(x, ex["x_ref_tol"]) < TOL


@pytest.mark.parametrize("name", GOLDEN)
def test_vcycle_matches_oracle_on_fresh_rhs(name, load_golden):
    """Same hierarchy, a different seeded rhs and a column-vector b: engine vs oracle."""
    ml, _ = load_golden(name)
    n = ml.levels[0].A.shape[0]
    b = np.random.default_rng(977).standard_normal((n, 1))
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.P)
    xo = cyc.solve(b, tol=0, maxiter=4)
    xg = ml.solve(b, tol=0, maxiter=4)
    assert xg.shape == (n,)          # the reference ravels b and x (multilevel.py:553-554) and returns (n,)
    assert relerr(xg, xo) < TOL


@pytest.mark.parametrize("env", [{"AMGB_NO_TILES": "1"}, {"AMGB_NO_PERMUTE": "1"},
                                 {"AMGB_NO_TILES": "1", "AMGB_NO_PERMUTE": "1"}, {"AMGB_NO_GRAPH": "1"},
                                 {"AMGB_TILE_G": "4"}, {"AMGB_TILE_G": "32"}, {"AMGB_TILE_CFG": "0"}, {"AMGB_TILE_CFG": "1"},
                                 {"AMGB_TILE_CFG": "4"},
                                 {"AMGB_TILE_CFG": "4", "AMGB_NO_HINTS": "1"}, {"AMGB_NO_TAIL": "1"},
                                 {"AMGB_TAIL_NNZ": "600000"}, {"AMGB_TAIL_NNZ": "600000", "AMGB_TAIL_CLUSTER": "1"},
                                 {"AMGB_TAIL_NNZ": "600000", "AMGB_TAIL_CLUSTER": "4"}, {"AMGB_NO_PDL": "1"},
                                 {"AMGB_NO_TAIL": "1", "AMGB_NO_TILES": "1"}, {"AMGB_TAIL_NNZ": "600000", "AMGB_TAIL_SOLO_BYTES": "0"},
                                 # the persistent grid kernel for the coarse part (grid_kernel.cuh), incl. no solo steps
                                 # and every step solo; the TMA tile kernels forced onto the small golden operators
                                 {"AMGB_TAIL_GRID": "1"}, {"AMGB_TAIL_GRID": "1", "AMGB_TAIL_SOLO_BYTES": "0"},
                                 {"AMGB_TAIL_GRID": "1", "AMGB_TAIL_SOLO_BYTES": "1e12"},
                                 {"AMGB_TAIL_GRID": "1", "AMGB_TAIL_NNZ": "3000", "AMGB_TILE_MIN_NNZ": "0"},
                                 {"AMGB_NO_TAIL": "1", "AMGB_TILE_MIN_NNZ": "0"},
                                 {"AMGB_NO_TAIL": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_TILE_G": "2"},
                                 {"AMGB_NO_TAIL": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_TILE_G": "8", "AMGB_TILE_CFG": "4"},
                                 {"AMGB_NO_TAIL": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_TILE_CFG": "1"},
                                 # second tile geometry for the denser operators, flat gathers for R / coarse P only,
                                 # other lanes-per-row rules
                                 {"AMGB_NO_TAIL": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_TILE_DENSE_CFG": "7", "AMGB_TILE_DENSE_AVG": "4"},
                                 {"AMGB_NO_TAIL": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_TILE_DENSE_CFG": "7", "AMGB_TILE_LANE_ENTRIES": "3"},
                                 {"AMGB_NO_TAIL": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_TILE_FLAT": "2"},
                                 {"AMGB_NO_TAIL": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_TILE_FLAT": "2", "AMGB_TILE_DENSE_CFG": "7"}])
@pytest.mark.parametrize("name", GOLDEN)
def test_every_kernel_path_matches_reference_golden(name, env, monkeypatch):
    """The lanes-per-row kernels, the un-permuted layout, the un-graphed cycle, forced lane-group widths
    and geometries of the TMA tile kernel, and the cluster tail kernel (off / 1 CTA / 4 CTAs / default 16)
    must all reproduce the reference (the env is read per hierarchy)."""
    from pyamg_b200.hierarchy_io import load_hierarchy
    from conftest import golden_path
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ml, ex = load_hierarchy(golden_path(name))
    x = ml.solve(ex["b"], tol=0, maxiter=len(ex["residuals"]) - 1)
    assert relerr(x, ex["x_ref"]) < TOL
    assert relerr(ml.solve(ex["b"], tol=0, maxiter=2, cycle="W"), ex["x_ref_W"]) < TOL


def test_callback_residuals_and_psolve(load_golden):
    ml, ex = load_golden("cfg3_rs_mcgs_poisson3d")
    seen, res = [], []
    x = ml.solve(ex["b"], tol=0, maxiter=3, callback=lambda xk: seen.append(xk.copy()), residuals=res)
    assert len(seen) == 3 and len(res) == 4
    assert relerr(seen[-1], x) == 0.0
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.P)
    assert relerr(x, cyc.solve(ex["b"], tol=0, maxiter=3)) < TOL
    assert relerr(ml.psolve(ex["b"]), cyc.solve(ex["b"], tol=0, maxiter=1)) < TOL


def test_aspreconditioner_with_scipy_cg(load_golden):
    """pyamg/tests/test_multilevel.py:45-67: the cycle as a preconditioner for SciPy CG."""
    from scipy.sparse.linalg import cg
    ml, ex = load_golden("cfg2_sa_jacobi_poisson2d")
    A = ml.levels[0].A
    M = ml.aspreconditioner(cycle="V")
    x, info = cg(A, ex["b"], rtol=1e-10, maxiter=40, M=M)
    assert info == 0
    assert np.linalg.norm(ex["b"] - A @ x) < 1e-8 * np.linalg.norm(ex["b"])
    res = []
    x2 = ml.solve(ex["b"], tol=1e-10, maxiter=40, accel="cg", residuals=res)
    assert np.linalg.norm(ex["b"] - A @ x2) < 1e-8 * np.linalg.norm(ex["b"])
    assert len(res) >= 2 and res[-1] < res[0]


def test_single_level_hierarchy_is_a_coarse_solve():
    """multilevel.py:559-561."""
    A = poisson1d(7)
    lvl = pyamg_b200.MultilevelSolver.Level()
    lvl.A = A
    ml = pyamg_b200.MultilevelSolver([lvl])
    b = np.arange(7, dtype=float)
    x = ml.solve(b, maxiter=1, tol=0)
    assert relerr(x, np.linalg.pinv(A.toarray()) @ b) < TOL


def test_errors_mirror_the_reference(load_golden):
    ml, ex = load_golden("cfg1_rs_gs_poisson2d")
    with pytest.raises(TypeError):
        ml.solve(ex["b"], cycle="Q")                      # multilevel.py:658
    with pytest.raises(ValueError):
        ml.solve(ex["b"], cycle="AMLI", accel="cg")       # multilevel.py:487-490
    with pytest.raises(ValueError):
        ml.solve(ex["b"][:-1])


# ------------------------------------------------------------------ single kernels vs golden / oracle
@pytest.mark.parametrize("name", GOLDEN)
def test_relaxation_kernels_match_reference_golden(name, load_golden):
    ml, ex = load_golden(name)
    lvl = ml.levels[0]
    A, x, b = lvl.A, ex["k_x"], ex["k_b"]
    for which, key in (("presmoother", "k_presmoother"), ("postsmoother", "k_postsmoother")):
        y = x.copy()
        getattr(lvl, which)(A, y, b)                      # the closures call the GPU shims
        assert relerr(y, ex[key]) < TOL
    Ac = A.tocsr()
    y = x.copy(); gpu_relax.jacobi(Ac, y, b, omega=0.7)
    assert relerr(y, ex["k_jacobi_w07"]) < TOL
    y = x.copy(); gpu_relax.gauss_seidel(Ac, y, b, sweep="symmetric")
    assert relerr(y, ex["k_gs_symmetric"]) < TOL
    y = x.copy(); gpu_relax.gauss_seidel(Ac, y, b, iterations=2, sweep="backward")
    assert relerr(y, ex["k_gs_backward2"]) < TOL
    y = x.copy(); gpu_relax.sor(Ac, y, b, omega=1.3)
    assert relerr(y, ex["k_sor_13"]) < TOL


@pytest.mark.parametrize("name", GOLDEN)
def test_matvecs_match_scipy_golden(name, load_golden):
    import ctypes
    from pyamg_b200 import _engine as E
    ml, ex = load_golden(name)
    lvl = ml.levels[0]
    for M, xin, key in ((lvl.A, ex["k_x"], "k_Ax"), (lvl.R, ex["k_x"], "k_Rx"), (lvl.P, ex["k_xc"], "k_Pxc")):
        keep = []
        Mc = E.as_matrix(M, keep)
        y = np.empty(M.shape[0])
        xin = np.ascontiguousarray(xin)
        E.check(E.lib().amgb_host_matvec(ctypes.byref(Mc), E.f64p(xin), E.f64p(y)))
        assert relerr(y, ex[key]) < TOL


def test_reference_kats_on_gpu():
    """The reference's own smoother KATs (tests/kats.py) against the CUDA kernels."""
    for case, got, exp in run_kats(gpu_relax):
        assert np.allclose(got, exp, rtol=1e-13, atol=1e-14), case
    r1, r2 = gs_convergence_case(gpu_relax)
    assert r1 < 0.01 and r2 < 0.01 and abs(r1 - r2) < 1e-7


def test_gs_lexicographic_equals_dense_triangular_solve():
    rng = np.random.default_rng(20260922)
    A = poisson1d(40).toarray() + 0.1 * np.diag(rng.random(40))
    As = sp.csr_array(A)
    for sweep in ("forward", "backward", "symmetric"):
        x = rng.random(40)
        b = rng.random(40)
        g = dense_gs_gold(A, x, b, sweep)
        gpu_relax.gauss_seidel(As, x, b, sweep=sweep)
        assert np.allclose(x, g, rtol=1e-12, atol=1e-13)


def test_bsr_jacobi_equals_csr_and_block_jacobi_gold():
    rng = np.random.default_rng(7919)
    N = 12
    A = poisson1d(N).toarray() + 0.3 * rng.random((N, N))
    for bs in (1, 2, 3, 4, 6):
        x0, b = rng.random(N), rng.random(N)
        xc = x0.copy(); oracle.jacobi(sp.csr_array(A), xc, b, omega=0.8)
        xb = x0.copy(); gpu_relax.jacobi(sp.bsr_array(A, blocksize=(bs, bs)), xb, b, omega=0.8)
        assert relerr(xb, xc) < TOL
        Dinv = np.stack([np.linalg.pinv(A[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs]) for i in range(N // bs)])
        for omega in (1.0, 1.1):
            x = x0.copy()
            g = block_jacobi_gold(A, x, b, Dinv, bs, omega)
            gpu_relax.block_jacobi(sp.bsr_array(A, blocksize=(bs, bs)), x, b, Dinv=Dinv, blocksize=bs, omega=omega)
            assert np.allclose(x, g, rtol=1e-12, atol=1e-13)


def test_make_system_contract_on_gpu_shims():
    """test_relaxation.py:48-111."""
    A = poisson1d(4)
    with pytest.raises(ValueError):
        gpu_relax.jacobi(A, [0, 0, 0, 0], np.zeros(4))
    with pytest.raises(ValueError):
        gpu_relax.jacobi(A, np.zeros(5), np.zeros(4))
    with pytest.raises(TypeError):
        gpu_relax.jacobi(A, np.zeros(4, dtype=np.float32), np.zeros(4))
    with pytest.raises(ValueError):
        gpu_relax.jacobi(A, np.zeros(8)[::2], np.zeros(4))
    with pytest.raises(ValueError):
        gpu_relax.gauss_seidel(A, np.zeros(4), np.zeros(4), sweep="diagonal")


def test_quirks_zero_diagonal_duplicates_unsorted_and_empty_rows():
    """SURVEY.md hazards 3: zero diagonal -> row untouched; LAST diagonal duplicate wins; unsorted
    column indices; empty rows; a ragged row much longer than the lane group."""
    rng = np.random.default_rng(31337)
    n = 257
    rows, cols, vals = [], [], []
    for i in range(n):
        if i % 17 == 3:
            continue                                   # empty row
        k = 200 if i == 100 else int(rng.integers(1, 9))
        cs = rng.choice(n, size=min(k, n), replace=False)
        for c in cs:
            if i % 11 == 5 and c == i:
                continue                               # no stored diagonal
            rows.append(i); cols.append(int(c)); vals.append(float(rng.standard_normal()))
        if i % 11 != 5:
            rows.append(i); cols.append(i); vals.append(4.0 + float(rng.random()))
        if i % 13 == 0 and i % 11 != 5:
            rows.append(i); cols.append(i); vals.append(9.0)   # duplicate diagonal: last one wins
    order = np.lexsort((rng.random(len(rows)), rows))          # row-major, columns shuffled
    rows, cols, vals = (np.array(v)[order] for v in (rows, cols, vals))
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    A = sp.csr_array((vals.astype(float), cols.astype(np.int32), indptr), shape=(n, n))
    x0, b = rng.random(n), rng.random(n)
    for fn, kw in (("jacobi", {"omega": 0.9, "iterations": 2}), ("gauss_seidel", {"sweep": "symmetric"}),
                   ("gauss_seidel", {"sweep": "backward", "omega": 0.8}),
                   ("gauss_seidel_indexed", {"indices": rng.permutation(n)[:200].astype(np.int32), "sweep": "forward"})):
        xo, xg = x0.copy(), x0.copy()
        getattr(oracle, fn)(A, xo, b, **kw)
        getattr(gpu_relax, fn)(A, xg, b, **kw)
        assert relerr(xg, xo) < TOL, fn


def test_linearity_and_larger_random_hierarchy_properties():
    """Size-independent properties: the V-cycle with x0=0 is a linear operator of b, and the
    iteration contracts the residual (convergence bound of classical/tests/test_classical.py:155-182
    style) on a 64^2 Poisson hierarchy built here."""
    from pyamg_b200.classical import ruge_stuben_solver
    from pyamg_b200.gallery import poisson
    A = poisson((64, 64))
    ml = ruge_stuben_solver(A, presmoother=("gauss_seidel_indexed", {"sweep": "symmetric"}),
                            postsmoother=("gauss_seidel_indexed", {"sweep": "symmetric"}))
    rng = np.random.default_rng(5)
    b1, b2 = rng.random(A.shape[0]), rng.random(A.shape[0])
    M = lambda v: ml.solve(v, tol=0, maxiter=1)
    assert relerr(M(2.0 * b1 - 3.0 * b2), 2.0 * M(b1) - 3.0 * M(b2)) < 1e-11
    res = []
    ml.solve(b1, tol=1e-10, maxiter=30, residuals=res)
    assert res[-1] < 1e-10 * np.linalg.norm(b1)
    assert (res[-1] / res[0]) ** (1.0 / (len(res) - 1)) < 0.3
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.P)
    assert relerr(ml.solve(b1, tol=0, maxiter=3), cyc.solve(b1, tol=0, maxiter=3)) < TOL


@pytest.mark.parametrize("name,n_dist", [("cfg3_rs_mcgs_poisson3d", 2), ("cfg4_sa_jacobi_aniso2d", 3),
                                         ("cfg1_rs_gs_poisson2d", 1)])
def test_distributed_layer_single_rank_on_gpu(name, n_dist, load_golden):
    if os.environ.get("AMGB_TEST_EMU") == "1":
        pytest.skip("the multi-GPU layer keeps its vectors in torch CUDA tensors: not available on the kernel emulator")
    """pyamg_b200.dist with the GPU backend at world_size 1 (tile kernels through amgb_operator_*, local
    wave-major order, replicated sub-hierarchy): same iterates as the oracle."""
    from pyamg_b200.dist import DistributedSolver, GpuBackend
    ml, ex = load_golden(name)
    be = GpuBackend(device=0, rank=0, world=1)
    ds = DistributedSolver(ml, be, n_dist=n_dist)
    ds.load(ex["b"])
    norms = be.vector(5)
    ds.cycles(4, norms=norms)
    x = ds.gather_x()
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.P)
    res = []
    xo = cyc.solve(ex["b"], tol=0, maxiter=4, residuals=res)
    assert relerr(x, xo) < TOL
    assert np.allclose(np.sqrt(norms[:5].cpu().numpy()), res, rtol=1e-9)
    be.close()


@pytest.mark.parametrize("name", GOLDEN)
def test_gpu_resident_pcg_matches_reference_golden(name, load_golden):
    """solve(accel='cg') with every vector in HBM (amgb_solve_cg) against the real reference's
    ml.solve(accel='cg'): same iteration count, info flag, residual history and iterate."""
    import warnings
    ml, ex = load_golden(name)
    if "x_ref_cg" not in ex:
        pytest.skip("no CG golden for this hierarchy (nonsymmetric operator)")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # non-symmetric smoother pairs warn, exactly like the reference
        res = []
        x, info = ml.solve(ex["b"], tol=1e-10, maxiter=10, accel="cg", residuals=res, return_info=True)
        assert info == int(ex["info_cg"][0]) and len(res) == len(ex["residuals_cg"])
        assert relerr(x, ex["x_ref_cg"]) < 1e-11
        assert np.allclose(res, ex["residuals_cg"], rtol=1e-7, atol=1e-12 * ex["residuals_cg"][0])
        res = []
        x0 = ex["x0"].copy()
        x, info = ml.solve(ex["b"], x0=x0, tol=1e-3, maxiter=30, accel="cg", cycle="W", residuals=res, return_info=True)
        assert np.array_equal(x0, ex["x0"])
        assert info == int(ex["info_cgW"][0]) and len(res) == len(ex["residuals_cgW"])
        assert relerr(x, ex["x_ref_cgW"]) < 1e-11
        # callback path: pyamg's CG on the host with the GPU cycle as M -- same numbers
        seen = []
        xc, infoc = ml.solve(ex["b"], tol=1e-10, maxiter=10, accel="cg", callback=lambda xk: seen.append(1),
                             return_info=True)
        assert infoc == int(ex["info_cg"][0]) and len(seen) == len(ex["residuals_cg"]) - 1
        assert relerr(xc, ex["x_ref_cg"]) < 1e-11


def test_out_buffer_and_pinned_result(load_golden):
    """solve(out=...) writes into the caller's (page-locked) buffer and returns it; x0 semantics unchanged."""
    ml, ex = load_golden("cfg3_rs_mcgs_poisson3d")
    n = len(ex["b"])
    out = pyamg_b200.pinned_empty(n)
    out[:] = np.nan
    x = ml.solve(ex["b"], tol=0, maxiter=len(ex["residuals"]) - 1, out=out)
    assert np.shares_memory(x, out) and relerr(x, ex["x_ref"]) < TOL
    x0 = ex["x0"].copy()
    x = ml.solve(ex["b"], x0=x0, tol=1e-6, maxiter=50, out=out)
    assert np.array_equal(x0, ex["x0"]) and relerr(x, ex["x_ref_tol"]) < TOL
    with pytest.raises(ValueError):
        ml.solve(ex["b"], out=np.zeros(n - 1))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sa_jacobi", "sa_block_jacobi"])
def test_large_levels_take_the_tile_path_at_default_thresholds(kind):
    """The goldens are small (<= 2 305 rows), so at the DEFAULT tile threshold (1.5 M entries per launch) they run on the
    lanes-per-row kernels.  Here the level-0 operators exceed the threshold -- smoothed aggregation + weighted Jacobi on
    Poisson 700^2 (2.4 M entries) and + block Jacobi on elasticity 260^2 (BSR(2,2), 2.4 M entries) -- so the TMA tile
    kernels (OP_JACOBI with the fused residual, OP_RESID, the flat-gather restriction) and block_jacobi_kernel run with
    the shipped defaults; two V-cycles against the oracle driving the reference's compiled kernels."""
    from pyamg_b200.aggregation import smoothed_aggregation_solver
    from pyamg_b200.gallery import linear_elasticity, poisson
    np.random.seed(7)
    if kind == "sa_jacobi":
        sm = ("jacobi", {"omega": 4.0 / 3.0})
        ml = smoothed_aggregation_solver(poisson((700, 700)), presmoother=sm, postsmoother=sm)
    else:
        A, B = linear_elasticity((260, 260))
        ml = smoothed_aggregation_solver(A, B=B, presmoother="block_jacobi", postsmoother="block_jacobi")
    A0 = ml.levels[0].A
    assert A0.nnz >= 1_500_000
    b = np.random.default_rng(11).random(A0.shape[0])
    kern = "ref" if oracle.have_ref() else "oracle"
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.dense_operator(ml.levels[-1].A),
                       kernels=kern)
    res_o, res_g = [], []
    xo = cyc.solve(b, tol=0, maxiter=2, residuals=res_o)
    xg = ml.solve(b, tol=0, maxiter=2, residuals=res_g)
    assert relerr(xg, xo) < TOL
    assert np.allclose(res_g, res_o, rtol=1e-10)
