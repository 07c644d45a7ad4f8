"""Build the BASELINE configs[2] hierarchy and run a few V-cycles (target for ncu captures).

    ncu ... python tools/run_cycles.py --grid 256 --cycles 2
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_hierarchy, SEED   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--cycles", type=int, default=2)
a = ap.parse_args()
ml = build_hierarchy((a.grid,) * 3)
b = np.random.default_rng(SEED).random(ml.levels[0].A.shape[0])
res = []
ml.solve(b, tol=0, maxiter=a.cycles, residuals=res)
print("residuals", res, "launches", ml.last_launches())
