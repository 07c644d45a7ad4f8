"""Fine-level kernel microbenchmark (GPU): SpMV / residual / fused Jacobi(+r) / GS colour sweep on
gallery.poisson(grid), timed with CUDA events on torch's current stream, L2 flushed between reps
(inputs at 256^3 are larger than L2 anyway).  Prints one JSON line per kernel.

    python tools/microbench.py --grid 256 256 256 --lanes 8 --reps 20
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
ap.add_argument("--lanes", type=int, nargs="+", default=[4, 8, 16])
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--peak", type=float, default=None)
a = ap.parse_args()

peak = a.peak
if peak is None:
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0

A = poisson(tuple(a.grid))
n, nnz = A.shape[0], A.nnz
dev = torch.device("cuda:0")
Ap = torch.from_numpy(A.indptr.astype(np.int32)).to(dev)
Aj = torch.from_numpy(A.indices.astype(np.int32)).to(dev)
Ax = torch.from_numpy(A.data).to(dev)
rng = np.random.default_rng(1)
x = torch.from_numpy(rng.random(n)).to(dev)
b = torch.from_numpy(rng.random(n)).to(dev)
y = torch.empty_like(x)
r = torch.empty_like(x)
# red-black colouring of the 7-pt stencil: colour = parity of the index sum
idx = np.indices(tuple(a.grid)).reshape(len(a.grid), -1).sum(0) & 1
rows = [torch.from_numpy(np.nonzero(idx == c)[0].astype(np.int32)).to(dev) for c in (0, 1)]
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
L = E.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())


def timeit(fn, reps):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


base = 12 * nnz + 4 * (n + 1)
for lanes in a.lanes:
    kernels = {
        "spmv": (lambda: E.check(L.amgb_dev_csr_spmv(n, P(Ap), P(Aj), P(Ax), P(x), P(y), lanes, st)), base + 16 * n),
        "residual": (lambda: E.check(L.amgb_dev_csr_residual(n, P(Ap), P(Aj), P(Ax), P(x), P(b), P(r), None, None, lanes, st)), base + 24 * n),
        "jacobi": (lambda: E.check(L.amgb_dev_csr_jacobi(n, P(Ap), P(Aj), P(Ax), P(x), P(b), P(y), None, 0.8, lanes, st)), base + 24 * n),
        "jacobi+r": (lambda: E.check(L.amgb_dev_csr_jacobi(n, P(Ap), P(Aj), P(Ax), P(x), P(b), P(y), P(r), 0.8, lanes, st)), base + 32 * n),
        "gs_2colour_sweep": (lambda: [E.check(L.amgb_dev_csr_gs_wave(len(rw), 0, P(rw), P(Ap), P(Aj), P(Ax), P(y), P(b), 1.0, lanes, st)) for rw in rows], base + 4 * n + 24 * n),
    }
    for name, (fn, nbytes) in kernels.items():
        med, best = timeit(fn, a.reps)
        print(json.dumps({"kernel": name, "grid": a.grid, "lanes": lanes, "ms_median": round(med, 4),
                          "ms_min": round(best, 4), "alg_GB": round(nbytes / 1e9, 4),
                          "GBps": round(nbytes / med / 1e6, 1), "frac_of_measured_peak": round(nbytes / med / 1e6 / peak, 3)}))
