"""Sweep the tile-kernel geometry / residency / L2-hint settings on the BASELINE configs[2] hierarchy.

    python tools/tune_tiles.py --grid 256
Prints, per setting, the graphed cycle time and the per-(level, op) achieved GB/s of the big levels.
"""
import argparse
import ctypes
import itertools
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_hierarchy, SEED, OPS   # noqa: E402
from pyamg_b200 import _engine as E            # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--settings", type=str, default="")
a = ap.parse_args()

ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
ml = build_hierarchy((a.grid,) * 3, stream=ts.cuda_stream)
n = ml.levels[0].A.shape[0]
b = torch.from_numpy(np.random.default_rng(SEED).random(n)).cuda()
x = torch.zeros_like(b)
P = lambda t: ctypes.c_void_p(t.data_ptr())
L = E.lib()

settings = [dict(zip(("AMGB_TILE_CFG", "AMGB_TILE_CTAS", "AMGB_NO_HINTS"), v)) for v in
            [("6", "9", "0"), ("4", "9", "0"), ("6", "6", "0")]]
if a.settings:      # "K=V,K=V;K=V" -> one dict of environment overrides per setting
    settings = [dict(kv.split("=") for kv in s.split(",") if kv) for s in a.settings.split(";")]
ALL_KEYS = sorted({k for st in settings for k in st})

x_first = None
for st in settings:
    for k in ALL_KEYS:
        os.environ.pop(k, None)
    os.environ.update(st)
    ml._invalidate()
    t0 = time.time()
    ml.upload()
    h = ml.handle
    x.zero_()
    E.check(L.amgb_solve_device(h, P(b), P(x), 3, 0, 1, None))
    torch.cuda.synchronize()
    if x_first is None:
        x_first = x.clone()
    dx = float((x - x_first).norm() / x_first.norm())     # same 3 cycles from zero under every setting
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    E.check(L.amgb_solve_device(h, P(b), P(x), 5, 0, 1, None))
    e1.record()
    torch.cuda.synchronize()
    cyc_ms = e0.elapsed_time(e1) / 5
    ml.profile_cycle()
    recs = np.concatenate([ml.profile_cycle() for _ in range(2)])
    g = {}
    for lvl, op, rows, nnz, nbytes, t in recs:
        k = (int(lvl), int(op))
        v = g.setdefault(k, [0.0, 0.0])
        v[0] += nbytes; v[1] += t
    line = {f"L{k[0]}:{OPS[k[1]].split('(')[0]}": round(v[0] / v[1] / 1e6) for k, v in sorted(g.items()) if k[0] <= 2}
    ms = {f"L{k[0]}:{OPS[k[1]].split('(')[0]}": round(v[1] / 2, 3) for k, v in sorted(g.items()) if k[0] <= 2}
    small = sum(v[1] for k, v in g.items() if k[0] >= 3) / 2
    print(json.dumps({"setting": st, "cycle_ms": round(cyc_ms, 3), "small_levels_ms": round(small, 3),
                      "relerr_vs_first_setting": dx, "GBps": line, "ms": ms}), flush=True)
