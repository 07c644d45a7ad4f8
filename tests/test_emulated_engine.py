"""CPU-side execution of the CUDA engine's own kernel sources (tests/emu): a bounded subset of the `gpu`
parity tests, run against libpyamg_b200_emu.so in a child process with AMGB_TEST_EMU=1.

What this proves: the LOGIC of the sm_100a code paths (TMA tile staging and offsets, mbarrier phases, wave
schedules, graph capture with baked pointers, the cluster tail interpreter, every fused epilogue, the new
8(f)-2 smoothers) reproduces the reference's goldens -- on the GPU-less build container, every round.
What it does not: performance, memory ordering, hardware limits.  The B200 run of `pytest -m gpu` remains the
parity gate; the product never loads the emulation library (tests/emu/cuda_runtime.h).

    AMGB_TEST_EMU=1 python -m pytest tests -m gpu -q          # the whole gpu suite on the emulator (~10 min)
    AMGB_TEST_EMU=1 AMGB_EMU_TMA=lazy ...                     # bulk copies complete at the wait, not at issue
"""
import os
import platform
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.skipif(platform.machine() != "x86_64" or shutil.which("g++") is None,
                                reason="the fiber emulator needs x86-64 and g++")

# fast, broad: every golden's V/W/F cycles, single-kernel goldens, quirks, KATs, PCG, a few forced kernel paths
SUBSET = ("vcycle_matches_reference_golden or w_and_f or relaxation_kernels or matvecs_match or quirks "
          "or reference_kats or pcg_matches or single_level or polynomial_matches or jacobi_indexed_and "
          "or block_gauss_seidel_matches or schwarz_matches or normal_equation_smoothers_match or device_mis "
          # round 2: the persistent grid kernel (env17 / env20), forced tile paths (env21), the second tile geometry and
          # flat gathers for R / coarse P (env25 / env28)
          "or (every_kernel_path and cfg3 and (env0 or env4 or env6 or env11 or env17 or env20 or env21 or env25 or env28))")


def _run(extra_env, k, files=("tests/test_gpu_parity.py", "tests/test_zz_gpu_widening.py"), timeout=900):
    env = dict(os.environ)
    env.update({"AMGB_TEST_EMU": "1"})
    env.update(extra_env)
    cmd = [sys.executable, "-m", "pytest", *files, "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "-k", k]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in r.stdout, tail
    return r.stdout


def test_emulator_reports_no_device_unless_asked():
    """The emulation library is not a fallback: without AMGB_TEST_EMU it has no device."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    lib = ctypes.CDLL(build_emu.build())
    saved = os.environ.pop("AMGB_TEST_EMU", None)
    try:
        assert lib.amgb_device_count() == 0
        h = ctypes.c_void_p()
        assert lib.amgb_hierarchy_create(0, ctypes.byref(h)) != 0
    finally:
        if saved is not None:
            os.environ["AMGB_TEST_EMU"] = saved


def test_gpu_parity_subset_on_the_emulator():
    out = _run({}, SUBSET)
    assert "failed" not in out


def test_tile_kernels_everywhere_with_lazy_tma_completion():
    """Every operator through the TMA tile kernels (AMGB_TILE_MIN_NNZ=0), bulk copies landing only when a
    barrier is waited on: a read of staged data before its wait would see poisoned shared memory."""
    _run({"AMGB_EMU_TMA": "lazy", "AMGB_TILE_MIN_NNZ": "0"},
         "vcycle_matches_reference_golden or relaxation_kernels or w_and_f")
    _run({"AMGB_EMU_TMA": "lazy", "AMGB_TILE_MIN_NNZ": "0", "AMGB_TILE_CFG": "1"}, "vcycle_matches_reference_golden",
         files=("tests/test_gpu_parity.py",))


def test_results_do_not_depend_on_thread_interleaving(tmp_path):
    """The emulator runs the lanes of a warp one after the other between synchronisation points; visiting warps and
    lanes in the opposite order or a pseudo-random one (AMGB_EMU_ORDER=reverse | random:<seed>) must not change a single bit of V/W cycles, GMRES or CG --
    with the lanes-per-row kernels and with the TMA tile kernels on every operator.  A kernel with a race (e.g. an
    in-place Gauss-Seidel wave that read a row of its own wave) would differ."""
    import numpy as np
    names = ["cfg3_rs_mcgs_poisson3d", "cfg5_sa_bjacobi_elasticity", "cfg9_air_fcjacobi_advection2d",
             "cfg10_sa_bgs_elasticity"]
    probe = os.path.join(ROOT, "tests", "emu", "order_probe.py")
    for extra in ({}, {"AMGB_TILE_MIN_NNZ": "0"}):
        dumps = []
        for order in ("forward", "reverse", "random:20260922"):
            env = dict(os.environ)
            env.update(extra)
            env.update({"AMGB_TEST_EMU": "1", "AMGB_EMU_ORDER": order})
            path = str(tmp_path / f"{order}{len(extra)}.npz")
            r = subprocess.run([sys.executable, probe, path, *names], cwd=ROOT, env=env, capture_output=True,
                               text=True, timeout=900)
            assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
            dumps.append(np.load(path))
        for n in names:
            for other in dumps[1:]:
                assert np.array_equal(dumps[0][n].view(np.int64), other[n].view(np.int64)), (n, extra)


@pytest.mark.parametrize("name,world,n_dist,halo", [("cfg3_rs_mcgs_poisson3d", 2, 2, "allgather"),
                                                    ("cfg4_sa_jacobi_aniso2d", 3, 2, "p2p"),
                                                    # halo exchange as ONE kernel storing into the neighbours' memory
                                                    # (amgb_comm_*; POSIX shared memory stands in for CUDA IPC)
                                                    ("cfg3_rs_mcgs_poisson3d", 2, 2, "peer"),
                                                    ("cfg4_sa_jacobi_aniso2d", 3, 2, "peer"),
                                                    ("cfg7_sa_cheby_richardson_poisson2d", 2, 2, "allgather")])
def test_distributed_cycle_on_the_emulator(name, world, n_dist, halo, tmp_path, load_golden):
    """The multi-GPU layer end to end at world_size 2 / 3: DistributedSolver + GpuBackend code paths with the
    emulated engine kernels, gloo collectives and CPU tensors standing in for device memory -- partitioned fine
    levels (halo all-gather or neighbour send/recv), all-reduced restriction, replicated coarse sub-hierarchy
    (amgb_solve_device) -- against the sequential oracle."""
    import socket
    import numpy as np
    import oracle
    from conftest import relerr
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "dist.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "emu", "dist_emu_worker.py"), name, str(n_dist), out, halo]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    got = np.load(out)
    ml, ex = load_golden(name)
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.P)
    res = []
    x = cyc.solve(ex["b"], tol=0, maxiter=3, residuals=res)
    assert relerr(got["x"], x) < 1e-12
    assert np.allclose(got["res"], res, rtol=1e-9)
    assert got["launches"][0] > 0


def test_persistent_grids_of_other_sizes():
    """The tile kernels run a persistent grid of (SM count x resident CTAs) blocks with a static tile -> warp map:
    with one emulated SM every warp walks many tiles, with 13 most warps get none -- same results either way."""
    for sms in ("1", "13"):
        _run({"AMGB_EMU_SMS": sms, "AMGB_TILE_MIN_NNZ": "0"},
             "vcycle_matches_reference_golden or relaxation_kernels", files=("tests/test_gpu_parity.py",))


@pytest.mark.skipif(os.environ.get("AMGB_TEST_ASAN") != "1", reason="opt-in (AMGB_TEST_ASAN=1): builds a second emulator library")
def test_gpu_subset_under_address_sanitizer():
    """The emulation library built with -fsanitize=address (AMGB_EMU_SANITIZE=1): device allocations are host heap
    blocks, so a kernel that reads or writes global memory out of bounds is reported.  The whole gpu suite runs the
    same way:
        LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
        AMGB_EMU_SANITIZE=1 AMGB_TEST_EMU=1 python -m pytest tests -m gpu -q"""
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not installed")
    _run({"LD_PRELOAD": asan, "ASAN_OPTIONS": "detect_leaks=0:detect_stack_use_after_return=0", "AMGB_EMU_SANITIZE": "1"},
         SUBSET + " or schwarz or normal_equation or spgemm", timeout=1800)
