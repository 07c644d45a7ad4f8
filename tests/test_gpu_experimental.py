"""Opt-in code paths written without GPU time left to validate them (end of round 1).  NOT selected by
`-m gpu` or `-m "not gpu"` runs of the default suites' GPU part: run with `-m gpu_experimental` on a B200
before enabling any of them by default:

    AMGB_EXPERIMENTAL=1 python -m pytest tests -m gpu_experimental -q


  AMGB_RESIDENT=1        DSMEM-resident Gauss-Seidel smoother applications (csrc/resident_kernel.cuh);
                         AMGB_RESIDENT_MAX_ROWS (default 65536) bounds the levels it takes
  AMGB_TILE_FLAT=1       flat-gather tile kernel (csrc/tile_flat_kernel.cuh): one gather round per tile at full
                         warp width, products reduced per row out of shared memory
  AMGB_TILE_PDL=1        programmatic dependent launch of the TMA tile kernel: the first operator tile is
                         requested before griddepcontrol.wait (csrc/tile_kernels.cuh, PDL = true)
  DistributedSolver.capture_graph / halo='p2p' on NCCL (pyamg_b200/dist.py)
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr, golden_path

pytestmark = [pytest.mark.gpu_experimental,
              pytest.mark.skipif(os.environ.get("AMGB_EXPERIMENTAL") != "1",
                                 reason="opt-in: AMGB_EXPERIMENTAL=1 pytest -m gpu_experimental (needs a B200)")]


@pytest.mark.parametrize("name", GOLDEN)
def test_resident_cluster_sweeps_match_reference_golden(name, monkeypatch):
    from pyamg_b200.hierarchy_io import load_hierarchy
    monkeypatch.setenv("AMGB_RESIDENT", "1")
    ml, ex = load_hierarchy(golden_path(name))
    x = ml.solve(ex["b"], tol=0, maxiter=len(ex["residuals"]) - 1)
    assert relerr(x, ex["x_ref"]) < 1e-12
    assert relerr(ml.solve(ex["b"], tol=0, maxiter=2, cycle="W"), ex["x_ref_W"]) < 1e-12
    rec = ml.profile_cycle()
    assert rec.shape[1] == 6


def test_resident_sweeps_on_a_mid_size_hierarchy(monkeypatch):
    import oracle
    from pyamg_b200.classical import ruge_stuben_solver
    from pyamg_b200.gallery import poisson
    monkeypatch.setenv("AMGB_RESIDENT", "1")
    sm = ("gauss_seidel_indexed", {"sweep": "symmetric"})
    ml = ruge_stuben_solver(poisson((48, 48, 48)), presmoother=sm, postsmoother=sm)
    b = np.random.default_rng(11).random(ml.levels[0].A.shape[0])
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.dense_operator(ml.levels[-1].A))
    assert relerr(ml.solve(b, tol=0, maxiter=3), cyc.solve(b, tol=0, maxiter=3)) < 1e-12


@pytest.mark.parametrize("env", [{"AMGB_TILE_PDL": "1", "AMGB_TILE_MIN_NNZ": "0"},
                                 {"AMGB_TILE_PDL": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_NO_GRAPH": "1"},
                                 {"AMGB_TILE_PDL": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_NO_PDL": "1"},
                                 {"AMGB_TILE_PDL": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_NO_HINTS": "1"},
                                 {"AMGB_TILE_PDL": "1"},
                                 {"AMGB_TILE_FLAT": "1", "AMGB_TILE_MIN_NNZ": "0"},
                                 {"AMGB_TILE_FLAT": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_NO_PERMUTE": "1"},
                                 {"AMGB_TILE_FLAT": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_NO_HINTS": "1", "AMGB_NO_GRAPH": "1"},
                                 {"AMGB_TILE_FLAT": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_TILE_PDL": "1"},
                                 {"AMGB_TILE_FLAT": "1"},
                                 {"AMGB_RESIDENT": "1", "AMGB_RESIDENT_MAX_ROWS": "1000000", "AMGB_TILE_PDL": "1"}])
@pytest.mark.parametrize("name", GOLDEN)
def test_tile_kernel_programmatic_launch_matches_reference_golden(name, env, monkeypatch):
    """Tile kernels launched with the programmatic-stream-serialization attribute, mixed with the rows
    kernel (with and without its own PDL), graphed and un-graphed."""
    from pyamg_b200.hierarchy_io import load_hierarchy
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ml, ex = load_hierarchy(golden_path(name))
    for _ in range(2):      # twice: the second solve replays the captured graph
        x = ml.solve(ex["b"], tol=0, maxiter=len(ex["residuals"]) - 1)
        assert relerr(x, ex["x_ref"]) < 1e-12
    assert relerr(ml.solve(ex["b"], tol=0, maxiter=2, cycle="F"), ex["x_ref_F"]) < 1e-12
    xt, info = ml.solve(ex["b"], x0=ex["x0"], tol=1e-6, maxiter=50, return_info=True)
    assert relerr(xt, ex["x_ref_tol"]) < 1e-12 and info == int(ex["info_tol"][0])


def test_tile_pdl_and_resident_on_a_mid_size_hierarchy(monkeypatch):
    import oracle
    from pyamg_b200.classical import ruge_stuben_solver
    from pyamg_b200.gallery import poisson
    monkeypatch.setenv("AMGB_TILE_PDL", "1")
    monkeypatch.setenv("AMGB_TILE_MIN_NNZ", "100000")
    monkeypatch.setenv("AMGB_RESIDENT", "1")
    sm = ("gauss_seidel_indexed", {"sweep": "symmetric"})
    ml = ruge_stuben_solver(poisson((64, 64, 64)), presmoother=sm, postsmoother=sm)
    b = np.random.default_rng(12).random(ml.levels[0].A.shape[0])
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.dense_operator(ml.levels[-1].A))
    assert relerr(ml.solve(b, tol=0, maxiter=3), cyc.solve(b, tol=0, maxiter=3)) < 1e-12


def test_flat_tile_kernel_on_a_mid_size_hierarchy_and_host_abi(monkeypatch):
    """Flat-gather tiles on operators with 7-60 entries per row (every epilogue), plus the smoother KATs
    through the host ABI (which builds one-operator hierarchies with the same kernels)."""
    import oracle
    from pyamg_b200.classical import ruge_stuben_solver
    from pyamg_b200.aggregation import smoothed_aggregation_solver
    from pyamg_b200.gallery import poisson
    monkeypatch.setenv("AMGB_TILE_FLAT", "1")
    monkeypatch.setenv("AMGB_TILE_MIN_NNZ", "0")
    sm = ("gauss_seidel_indexed", {"sweep": "symmetric"})
    ml = ruge_stuben_solver(poisson((40, 40, 40)), presmoother=sm, postsmoother=sm)
    b = np.random.default_rng(13).random(ml.levels[0].A.shape[0])
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.dense_operator(ml.levels[-1].A))
    assert relerr(ml.solve(b, tol=0, maxiter=3), cyc.solve(b, tol=0, maxiter=3)) < 1e-12
    jac = ("jacobi", {"omega": 4.0 / 3.0, "iterations": 2})
    ml = smoothed_aggregation_solver(poisson((300, 300)), presmoother=jac, postsmoother=jac)
    b = np.random.default_rng(14).random(ml.levels[0].A.shape[0])
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.dense_operator(ml.levels[-1].A))
    assert relerr(ml.solve(b, tol=0, maxiter=3, cycle="W"), cyc.solve(b, tol=0, maxiter=3, cycle="W")) < 1e-12


@pytest.mark.parametrize("case", range(6))
def test_experimental_kernels_on_random_hierarchies(case, monkeypatch):
    """The seeded random two-level hierarchies of tests/test_zz_gpu_widening.py (rows longer than a tile, zero
    diagonal, unsorted indices, random smoothers / layouts) through the opt-in kernel variants."""
    from test_zz_gpu_widening import random_hierarchy_check
    random_hierarchy_check(case, monkeypatch, (
        {"AMGB_TILE_FLAT": "1", "AMGB_TILE_MIN_NNZ": "0"},
        {"AMGB_RESIDENT": "1", "AMGB_RESIDENT_MAX_ROWS": "100000"},
        {"AMGB_TILE_PDL": "1", "AMGB_TILE_MIN_NNZ": "0"},
        {"AMGB_TILE_FLAT": "1", "AMGB_TILE_MIN_NNZ": "0", "AMGB_RESIDENT": "1", "AMGB_TILE_PDL": "1"}))
