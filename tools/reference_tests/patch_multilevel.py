"""Route the reference's MultilevelSolver.solve / aspreconditioner through pyamg_b200 and
run the reference's OWN tests: setup stays on the reference, every solve runs on the engine (kernel emulator here)."""
import os, sys
ROOT = os.environ.get("AMGB_REPO_ROOT") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
os.environ["AMGB_TEST_EMU"] = "1"
import conftest as _amgb_conftest
import numpy as np
import pyamg.multilevel as M
import pyamg_b200
_orig_solve, _orig_asprec = M.MultilevelSolver.solve, M.MultilevelSolver.aspreconditioner
STATS = {"gpu": 0, "fallback": 0}

def _gpu(self):
    g = getattr(self, "_gpu", None)
    if g is None or getattr(self, "_gpu_levels", None) != [id(l.A) for l in self.levels]:
        g = pyamg_b200.MultilevelSolver.from_pyamg(self)
        self._gpu, self._gpu_levels = g, [id(l.A) for l in self.levels]
    return g

def solve(self, b, *a, **kw):
    A = self.levels[0].A
    if np.iscomplexobj(A.data) or A.dtype != np.float64 or np.iscomplexobj(b) or np.asarray(b).dtype != np.float64:
        STATS["fallback"] += 1
        return _orig_solve(self, b, *a, **kw)        # complex / f32: outside the engine's scope
    STATS["gpu"] += 1
    return _gpu(self).solve(b, *a, **kw)

M.MultilevelSolver.solve = solve
import atexit
atexit.register(lambda: print("\n[patch_ml_plugin] solves on the engine: %d, left to the reference (complex / f32): %d" % (STATS["gpu"], STATS["fallback"])))
