"""pytest plugin: replace the reference's relaxation functions by pyamg_b200's GPU mirrors (kernel emulator in the
build container) so that the reference's OWN test file pyamg/relaxation/tests/test_relaxation.py exercises them."""
import os, sys
ROOT = os.environ.get("AMGB_REPO_ROOT") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
os.environ["AMGB_TEST_EMU"] = "1"
import conftest as _amgb_conftest   # installs the emulator as the engine
import pyamg.relaxation.relaxation as R
import pyamg_b200.relaxation.relaxation as G
NAMES = ["jacobi", "gauss_seidel", "sor", "gauss_seidel_indexed", "block_jacobi", "block_gauss_seidel", "polynomial",
         "jacobi_indexed", "cf_jacobi", "fc_jacobi", "cf_block_jacobi", "fc_block_jacobi", "jacobi_ne", "gauss_seidel_ne",
         "gauss_seidel_nr", "schwarz"]
for n in NAMES:
    setattr(R, n, getattr(G, n))
import pyamg.relaxation as RP
for n in NAMES:
    if hasattr(RP, n):
        setattr(RP, n, getattr(G, n))
