"""Fine-level kernel microbenchmark (GPU): the TMA tile kernels through the resident-operator C API on
gallery.poisson(grid) in natural order: SpMV / residual / fused Jacobi (with and without the residual
by-product), CUDA events on a non-default torch stream; inputs at 256^3 (1.4 GB operator) exceed L2.

    python tools/microbench.py --grid 256 256 256 --reps 20
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyamg_b200 import _engine as E      # noqa: E402
from pyamg_b200.gallery import poisson   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, nargs="+", default=[256, 256, 256])
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0

ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
A = poisson(tuple(a.grid))
n, nnz = A.shape[0], A.nnz
L = E.lib()
keep = []
Ac = E.as_matrix(A, keep)
op = ctypes.c_void_p()
E.check(L.amgb_operator_create(0, ctypes.byref(Ac), None, 0, ctypes.c_void_p(ts.cuda_stream), ctypes.byref(op)))
rng = np.random.default_rng(1)
mk = lambda: torch.from_numpy(np.concatenate([rng.random(n), [0, 0]])).cuda()
x, b, y, r = mk(), mk(), mk(), mk()
n2 = torch.zeros(1, dtype=torch.float64, device="cuda")
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
base = 12 * nnz + 4 * (n + 1)
cases = {
    "spmv": (lambda: L.amgb_operator_apply(op, 0, P(x), None, P(y), None, 0.0, None, -1), base + 16 * n),
    "residual": (lambda: L.amgb_operator_apply(op, 1, P(x), P(b), P(r), None, 0.0, None, -1), base + 24 * n),
    "residual+norm": (lambda: L.amgb_operator_apply(op, 1, P(x), P(b), P(r), None, 0.0, P(n2), -1), base + 24 * n),
    "jacobi": (lambda: L.amgb_operator_apply(op, 3, P(x), P(b), P(y), None, 0.8, None, -1), base + 24 * n),
    "jacobi+residual(fused)": (lambda: L.amgb_operator_apply(op, 3, P(x), P(b), P(y), P(r), 0.8, None, -1), base + 32 * n),
}
for name, (fn, nbytes) in cases.items():
    for _ in range(3):
        E.check(fn())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        E.check(fn())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print(json.dumps({"kernel": name, "grid": a.grid, "ms": round(ms, 4), "alg_GB": round(nbytes / 1e9, 4),
                      "GBps": round(nbytes / ms / 1e6, 1), "frac_of_measured_peak": round(nbytes / ms / 1e6 / peak, 3)}), flush=True)
