"""pytest configuration: `gpu` marker + shared helpers.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol export (no compute calls).
`-m gpu`      : parity tests proper, through the C ABI of libpyamg_b200.so on a real B200.
"""
import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# The unmodified reference, installed by oracle/build.py into baseline/_ref (git-ignored, travels with the gpurun
# snapshot): with it `import pyamg` works on the GPU box too, so the from_pyamg adoption tests run there.
_REF_SITE = os.path.join(ROOT, "baseline", "_ref")
if os.path.exists(os.path.join(_REF_SITE, "pyamg", "__init__.py")) and _REF_SITE not in sys.path:
    sys.path.append(_REF_SITE)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
def _cfg_number(name):
    import re
    return int(re.match(r"cfg(\d+)", name).group(1))


GOLDEN_ALL = sorted((os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))), key=_cfg_number)
# cfg1-6: the five BASELINE families (+ descriptor coverage), validated on a B200 in round 1.
# cfg7+ : SURVEY 8(f)-2 widening (polynomial / CF-FC Jacobi / AIR / block Gauss-Seidel); their GPU tests live in
#         tests/test_zz_gpu_widening.py, which sorts last so that `pytest -x` reaches them after everything else.
GOLDEN = [g for g in GOLDEN_ALL if _cfg_number(g) <= 6]
GOLDEN_WIDENING = [g for g in GOLDEN_ALL if _cfg_number(g) > 6]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "timeout(seconds): pytest-timeout's per-test limit (a no-op without the plugin)")
    config.addinivalue_line("markers", "gpu_experimental: opt-in code paths not yet validated on hardware "
                                       "(run explicitly with -m gpu_experimental; never part of -m gpu)")


def _install_emulator():
    """AMGB_TEST_EMU=1: run the `gpu` tests against tests/emu (the engine's kernel sources executed on host
    fibers) instead of a B200.  Test infrastructure only -- the product never loads that library."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from pyamg_b200 import _engine as E
    E._lib = E._bind(ctypes.CDLL(build_emu.build()))


if os.environ.get("AMGB_TEST_EMU") == "1":
    _install_emulator()


def golden_path(name):
    return os.path.join(GOLDEN_DIR, name + ".npz")


@pytest.fixture(scope="session")
def load_golden():
    from pyamg_b200.hierarchy_io import load_hierarchy
    cache = {}

    def _load(name):
        if name not in cache:
            cache[name] = load_hierarchy(golden_path(name))
        return cache[name]

    return _load


def relerr(a, b):
    import numpy as np
    a = np.asarray(a, dtype=float).ravel()
    b = np.asarray(b, dtype=float).ravel()
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def summation_order_sensitivity(ml, b, **solve_kw):
    """How far the ORACLE's own result moves when every SpMV sums its row in the opposite order (a pure
    rounding-level perturbation).  The engine's only arithmetic difference to the reference is the summation
    order inside a row, so for ill-conditioned recurrences (AMLI's A-orthogonalised step sizes on nearly
    collinear corrections) this is the honest yardstick; for V/W/F cycles it is ~1e-16."""
    import numpy as np
    import scipy.sparse as sp
    import oracle
    cyc = oracle.Cycle(oracle.hierarchy_spec(ml), coarse_pinv=ml.coarse_solver.P)
    x = cyc.solve(b, **solve_kw)
    orig = oracle.matvec

    def reversed_rows(A, v, kernels="oracle"):
        C = A.tocsr()
        ip = C.indptr
        idx = (np.concatenate([np.arange(ip[i + 1] - 1, ip[i] - 1, -1) for i in range(C.shape[0])])
               if C.nnz else np.zeros(0, dtype=int))
        return orig(sp.csr_array((C.data[idx], C.indices[idx], ip), shape=C.shape), v, kernels)

    oracle.matvec = reversed_rows
    try:
        x2 = cyc.solve(b, **solve_kw)
    finally:
        oracle.matvec = orig
    return relerr(x2, x)
