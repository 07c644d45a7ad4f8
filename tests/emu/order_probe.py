"""Child process of tests/test_emulated_engine.py: solve a few goldens on the kernel emulator and dump the iterates.
Run twice (AMGB_EMU_ORDER unset / =reverse): bit-identical dumps mean no kernel result depends on how threads are
interleaved between synchronisation points."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
os.environ["AMGB_TEST_EMU"] = "1"
import conftest  # noqa: E402,F401  (installs the emulator)
from conftest import golden_path  # noqa: E402
from pyamg_b200.hierarchy_io import load_hierarchy  # noqa: E402

warnings.simplefilter("ignore")
out = {}
for name in sys.argv[2:]:
    ml, ex = load_hierarchy(golden_path(name))
    out[name] = np.concatenate([ml.solve(ex["b"], tol=0, maxiter=2),
                                ml.solve(ex["b"], tol=0, maxiter=1, cycle="W"),
                                ml.solve(ex["b"], tol=1e-8, maxiter=4, accel="gmres"),
                                ml.solve(ex["b"], tol=1e-8, maxiter=4, accel="cg")])
np.savez(sys.argv[1], **out)
