"""Opt-in code paths written without GPU time left to validate them (end of round 1).  NOT selected by
`-m gpu` or `-m "not gpu"` runs of the default suites' GPU part: run with `-m gpu_experimental` on a B200
before enabling any of them by default:

    AMGB_EXPERIMENTAL=1 python -m pytest tests -m gpu_experimental -q


  AMGB_RESIDENT=1        DSMEM-resident Gauss-Seidel smoother applications (csrc/resident_kernel.cuh)
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
