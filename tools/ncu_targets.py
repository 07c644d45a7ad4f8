"""Targets for the round-2 ncu captures (tools/r2_ncu.sh): every shipped hot kernel launched a few times in isolation
on BASELINE-sized operands, so one `ncu -k regex:<kernel>` pass per kernel family finds it without wading through a cycle.

    python tools/ncu_targets.py fine      # level-0 operator of poisson 256^3, natural order: OP 0 / 1 / 3(+r) tile kernels
    python tools/ncu_targets.py cycle     # one un-graphed V-cycle of the 256^3 RS hierarchy (AMGB_NO_GRAPH=1)
    python tools/ncu_targets.py cfg5      # one un-graphed V-cycle of elasticity 300^2 (block_jacobi_kernel)
    python tools/ncu_targets.py cfg2      # one un-graphed V-cycle of SA + Jacobi 2000^2
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                   # noqa: E402
from pyamg_b200 import _engine as E            # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "fine"
os.environ["AMGB_NO_GRAPH"] = "1"
if what == "fine":
    import torch
    from pyamg_b200.gallery import poisson
    A0 = poisson((256, 256, 256))
    n = A0.shape[0]
    L = E.lib()
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    keep = []
    A0c = E.as_matrix(A0, keep)
    op0 = ctypes.c_void_p()
    E.check(L.amgb_operator_create(0, ctypes.byref(A0c), None, 0, ctypes.c_void_p(ts.cuda_stream), ctypes.byref(op0)))
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n + 2, dtype=torch.float64, device="cuda", generator=g)
    b = torch.rand(n + 2, dtype=torch.float64, device="cuda", generator=g)
    y, r = torch.empty_like(x), torch.empty_like(x)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    for _ in range(2):
        E.check(L.amgb_operator_apply(op0, 0, P(x), None, P(y), None, 0.0, None, -1))      # SpMV
        E.check(L.amgb_operator_apply(op0, 1, P(x), P(b), P(y), None, 0.0, None, -1))      # residual
        E.check(L.amgb_operator_apply(op0, 3, P(x), P(b), P(y), P(r), 0.8, None, -1))      # fused Jacobi + residual
        E.check(L.amgb_operator_apply(op0, 3, P(x), P(b), P(y), None, 0.8, None, -1))      # Jacobi alone
    torch.cuda.synchronize()
    L.amgb_operator_destroy(op0)
else:
    bench.WORKLOAD["name"] = {"cycle": "cfg3"}.get(what, what)
    g = bench.DEFAULT_GRID[bench.WORKLOAD["name"]]
    ml = bench.build_hierarchy((g,) * 3)
    b = np.random.default_rng(bench.SEED).random(ml.levels[0].A.shape[0])
    res = []
    ml.solve(b, tol=0, maxiter=1, residuals=res)
    print("residuals", res, "launches", ml.last_launches())
