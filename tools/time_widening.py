"""Timing of the SURVEY 8(f) rows that were written without GPU time (round 1): run on a B200 in round 2.

    python tools/time_widening.py --grid 128          # ~2 M unknowns; prints one JSON line per item

  * Galerkin product R A P: GPU SpGEMM (amgb_host_csr_matmat, host buffers in/out) vs SciPy on the host, bitwise
    comparison of the results included;
  * rho(D^-1 A): restarted Arnoldi on the host vs amgb_arnoldi_* (basis resident in HBM);
  * smoothers: V-cycles/s of the same RS hierarchy with multi-colour GS, Chebyshev(3), CF-Jacobi;
  * Krylov: solve(accel='cg' | 'gmres' | 'fgmres') to 1e-8, iterations and wall time.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import scipy.sparse as sp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    args = ap.parse_args()
    from pyamg_b200 import _engine as E
    from pyamg_b200.classical import ruge_stuben_solver
    from pyamg_b200.gallery import poisson
    from pyamg_b200.relaxation.smoothing import change_smoothers
    g = args.grid
    A = poisson((g, g, g))
    t0 = time.perf_counter()
    ml = ruge_stuben_solver(A, presmoother=("gauss_seidel_indexed", {"sweep": "symmetric"}),
                            postsmoother=("gauss_seidel_indexed", {"sweep": "symmetric"}))
    print(json.dumps({"item": "host_setup_s", "grid": g, "value": round(time.perf_counter() - t0, 2),
                      "levels": [lv.A.shape[0] for lv in ml.levels]}), flush=True)
    # ---- Galerkin product
    for l, lvl in enumerate(ml.levels[:2]):
        R, Al, P = sp.csr_array(lvl.R), sp.csr_array(lvl.A), sp.csr_array(lvl.P)
        t0 = time.perf_counter()
        Cs = (R @ Al) @ P
        t_host = time.perf_counter() - t0
        E.csr_matmat(sp.eye(4, format="csr"), sp.eye(4, format="csr"))         # context / module warm-up
        t0 = time.perf_counter()
        Cg = E.csr_matmat(E.csr_matmat(R, Al), P)
        t_gpu = time.perf_counter() - t0
        same = (np.array_equal(Cs.indptr, Cg.indptr) and np.array_equal(Cs.indices, Cg.indices)
                and np.array_equal(Cs.data.view(np.int64), Cg.data.view(np.int64)))
        print(json.dumps({"item": "galerkin", "level": l, "n": Al.shape[0], "nnz_A": int(Al.nnz), "nnz_C": int(Cs.nnz),
                          "scipy_s": round(t_host, 3), "gpu_e2e_s": round(t_gpu, 3), "bitwise_equal": bool(same)}),
              flush=True)
    # ---- spectral-radius estimate rho(D^-1 A): host Arnoldi vs resident-basis Arnoldi on the device
    from pyamg_b200.util import approximate_spectral_radius, get_diagonal
    A0 = sp.csr_array(ml.levels[0].A)
    D = get_diagonal(A0, inv=True)
    t0 = time.perf_counter()
    rh = approximate_spectral_radius(A0, row_scale=D, where="host")
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    rg = approximate_spectral_radius(A0, row_scale=D, where="gpu")
    t_gpu = time.perf_counter() - t0
    print(json.dumps({"item": "rho_Dinv_A", "n": A0.shape[0], "host_s": round(t_host, 3), "gpu_e2e_s": round(t_gpu, 3),
                      "rel_diff": float(abs(rh - rg) / rh)}), flush=True)
    # ---- vertex colouring ('MIS' of the reference = natural-order first fit): host routine vs device wavefront
    from pyamg_b200.graph import vertex_coloring
    for l, lvl in enumerate(ml.levels[:3]):
        G = sp.csr_array(lvl.A)
        t0 = time.perf_counter()
        ch = vertex_coloring(G, "MIS", where="host")
        t_host = time.perf_counter() - t0
        vertex_coloring(sp.csr_array(ml.levels[-1].A), "MIS", where="gpu")       # context / module warm-up
        t0 = time.perf_counter()
        cg = vertex_coloring(G, "MIS", where="gpu")
        t_gpu = time.perf_counter() - t0
        print(json.dumps({"item": "coloring_mis", "level": l, "n": G.shape[0], "nnz": int(G.nnz),
                          "colors": int(ch.max()) + 1, "host_s": round(t_host, 3), "gpu_e2e_s": round(t_gpu, 3),
                          "identical": bool(np.array_equal(ch, cg))}), flush=True)
    # ---- smoothers
    n = A.shape[0]
    b = np.random.default_rng(20260922).random(n)
    for label, sm in (("mc_gs_symmetric", ("gauss_seidel_indexed", {"sweep": "symmetric"})),
                      ("chebyshev3", ("chebyshev", {"degree": 3})),
                      ("cf_jacobi", ("cf_jacobi", {"omega": 0.8, "f_iterations": 2})),
                      ("jacobi", ("jacobi", {"omega": 4.0 / 3.0}))):
        change_smoothers(ml, sm, sm)
        ml.solve(b, tol=0, maxiter=2)                                           # upload + graph capture
        res = []
        t0 = time.perf_counter()
        ml.solve(b, tol=0, maxiter=10, residuals=res)
        dt = time.perf_counter() - t0
        print(json.dumps({"item": "cycles", "smoother": label, "e2e_cycles_per_s": round(10 / dt, 1),
                          "residual_drop_per_cycle": round(float((res[-1] / res[0]) ** 0.1), 4)}), flush=True)
    # ---- Krylov
    change_smoothers(ml, ("gauss_seidel_indexed", {"sweep": "symmetric"}), ("gauss_seidel_indexed", {"sweep": "symmetric"}))
    for accel in (None, "cg", "gmres", "fgmres"):
        res = []
        ml.solve(b, tol=1e-8, maxiter=30, accel=accel, residuals=res)           # warm-up (buffers, graphs)
        res = []
        t0 = time.perf_counter()
        ml.solve(b, tol=1e-8, maxiter=30, accel=accel, residuals=res)
        print(json.dumps({"item": "solve_1e-8", "accel": accel, "iterations": len(res) - 1,
                          "seconds": round(time.perf_counter() - t0, 4)}), flush=True)
    # ---- smoothers whose waves come from conflict graphs (Kaczmarz, Schwarz): smaller problem, the Schwarz setup
    #      inverts one dense block per row on the host (as the reference does)
    gs = min(g, 48)
    ml = ruge_stuben_solver(poisson((gs, gs, gs)))
    b = np.random.default_rng(20260923).random(ml.levels[0].A.shape[0])
    for label, sm in (("gauss_seidel", ("gauss_seidel", {"sweep": "symmetric"})),
                      ("gauss_seidel_ne", ("gauss_seidel_ne", {"sweep": "symmetric"})),
                      ("gauss_seidel_nr", ("gauss_seidel_nr", {"sweep": "symmetric"})),
                      ("schwarz", ("schwarz", {"sweep": "symmetric"}))):
        t0 = time.perf_counter()
        change_smoothers(ml, sm, sm)
        ml.solve(b, tol=0, maxiter=2)
        t_setup = time.perf_counter() - t0
        res = []
        t0 = time.perf_counter()
        ml.solve(b, tol=0, maxiter=10, residuals=res)
        dt = time.perf_counter() - t0
        print(json.dumps({"item": "cycles_small", "grid": gs, "smoother": label, "setup_and_upload_s": round(t_setup, 2),
                          "e2e_cycles_per_s": round(10 / dt, 1),
                          "residual_drop_per_cycle": round(float((res[-1] / res[0]) ** 0.1), 4)}), flush=True)


if __name__ == "__main__":
    main()
