"""Save / load a multigrid hierarchy (operators + smoother descriptors) as one .npz file.

The reference has no persistence ("the hierarchy is just a Python list"); this is
the engine's resume path -- setup of the large BASELINE configs costs minutes of CPU, the upload
seconds -- and the container of the golden fixtures under tests/golden/.

Layout: ``meta`` (JSON: per level the operator formats/shapes/blocksizes and the smoother
descriptors ``{"fn": <relaxation function name>, "name": <registry __name__>, "kw": {...}}``),
arrays ``L<k>_<A|P|R>_{indptr,indices,data}``, ``L<k>_<pre|post>_<indices|Dinv>``, ``coarse_P``,
plus any extra arrays the caller adds (golden vectors).
"""
import json
from functools import partial, update_wrapper

import numpy as np
from scipy import sparse

from .multilevel import MultilevelSolver
from .relaxation import relaxation

_ARRAY_KW = ("indices", "Dinv", "Cpts", "Fpts", "coefficients")
_INT_KW = ("iterations", "blocksize", "f_iterations", "c_iterations")
_SCHWARZ_ARRAYS = ("subdomain", "subdomain_ptr", "inv_subblock", "inv_subblock_ptr")


def _put_matrix(out, key, M):
    if getattr(M, "format", None) not in ("csr", "bsr"):
        M = M.tocsr()
    out[key + "_indptr"] = np.asarray(M.indptr, dtype=np.int32)
    out[key + "_indices"] = np.asarray(M.indices, dtype=np.int32)
    out[key + "_data"] = np.asarray(M.data, dtype=np.float64)
    return {"format": M.format, "shape": [int(M.shape[0]), int(M.shape[1])],
            "blocksize": [int(v) for v in (M.blocksize if M.format == "bsr" else (1, 1))]}


def _get_matrix(z, key, m):
    arrs = (z[key + "_data"], z[key + "_indices"], z[key + "_indptr"])
    if m["format"] == "bsr":
        return sparse.bsr_array(arrs, shape=tuple(m["shape"]), blocksize=tuple(m["blocksize"]))
    return sparse.csr_array(arrs, shape=tuple(m["shape"]))


def _put_smoother(out, key, sm, A=None):
    if sm is None or (getattr(sm, "func", None) is None and getattr(sm, "__name__", "") == "none"):
        return None
    func = getattr(sm, "func", None)
    if func is None:
        from .relaxation.smoothing import polynomial_closure_parameters, normal_equation_closure_parameters
        ne = normal_equation_closure_parameters(sm)       # 'jacobi_ne' / 'gauss_seidel_ne' / 'gauss_seidel_nr' closures
        if ne is not None:
            kw = {k: (int(v) if k == "iterations" else (v if isinstance(v, str) else float(np.real(v)))) for k, v in ne[1].items()}
            return {"fn": ne[0], "name": ne[0], "kw": kw, "closure": "ne"}
        from .relaxation.smoothing import schwarz_closure_parameters
        sz = schwarz_closure_parameters(sm, A)            # 'schwarz' / 'strength_based_schwarz' closures
        if sz is not None:
            for k in _SCHWARZ_ARRAYS:
                out[f"{key}_{k}"] = np.asarray(sz[k])
            return {"fn": "schwarz", "name": "schwarz", "closure": "schwarz",
                    "kw": {"iterations": int(sz["iterations"]), "sweep": sz["sweep"]}}
        poly = polynomial_closure_parameters(sm)          # 'richardson' / 'chebyshev' closures
        if poly is None:
            raise NotImplementedError(f"cannot serialise closure smoother {sm!r}")
        out[f"{key}_coefficients"] = np.asarray(poly[0], dtype=np.float64)
        return {"fn": "polynomial", "name": sm.__name__, "kw": {"iterations": int(poly[1])}, "closure": True}
    kw = {}
    for k, v in sm.keywords.items():
        if k in _ARRAY_KW:
            out[f"{key}_{k}"] = np.asarray(v)
        elif isinstance(v, (str, bool)) or v is None:
            kw[k] = v
        elif np.isscalar(v) or (isinstance(v, np.ndarray) and v.size == 1):
            f = float(np.real(np.asarray(v).reshape(-1)[0]))
            kw[k] = int(f) if k in _INT_KW else f
        else:
            raise NotImplementedError(f"smoother keyword {k}={v!r}")
    return {"fn": func.__name__, "name": getattr(sm, "__name__", func.__name__), "kw": kw}


def _get_smoother(z, key, d, lvl=None):
    if d is not None and d.get("closure") == "ne":
        from .relaxation.smoothing import _ne_closure
        return _ne_closure(d["name"], lvl, **d["kw"])
    if d is not None and d.get("closure") == "schwarz":
        from .relaxation.smoothing import setup_schwarz
        return setup_schwarz(lvl, **d["kw"], **{k: z[f"{key}_{k}"] for k in _SCHWARZ_ARRAYS})
    if d is None:
        def none(A, x, b):
            pass
        return none
    fn = getattr(relaxation, d["fn"])
    kw = dict(d["kw"])
    for k in _ARRAY_KW:
        if f"{key}_{k}" in z:
            kw[k] = z[f"{key}_{k}"]
    if d.get("closure"):
        from .relaxation.smoothing import _polynomial_closure
        return _polynomial_closure(d["name"], kw["coefficients"], kw.get("iterations", 1))
    sm = partial(fn, **kw)
    update_wrapper(sm, getattr(relaxation, d["name"], fn))
    return sm


def save_hierarchy(path, ml, extra=None, compressed=True):
    """Write ``ml`` (a pyamg or pyamg_b200 MultilevelSolver) to ``path``; ``extra`` = more arrays."""
    out, meta = {}, {"levels": []}
    for k, lvl in enumerate(ml.levels):
        m = {"A": _put_matrix(out, f"L{k}_A", lvl.A)}
        if k < len(ml.levels) - 1:
            R = lvl.R if hasattr(lvl, "R") else lvl.P.T.conjugate()
            m["P"] = _put_matrix(out, f"L{k}_P", lvl.P)
            m["R"] = _put_matrix(out, f"L{k}_R", R)
            m["pre"] = _put_smoother(out, f"L{k}_pre", getattr(lvl, "presmoother", None), lvl.A)
            m["post"] = _put_smoother(out, f"L{k}_post", getattr(lvl, "postsmoother", None), lvl.A)
        meta["levels"].append(m)
    from .multilevel import coarse_solver_spec
    meta["coarse_solver"] = repr(coarse_solver_spec(ml.coarse_solver))     # name or (name, kwargs)
    meta["symmetric_smoothing"] = bool(getattr(ml, "symmetric_smoothing", False))
    cached = getattr(ml.coarse_solver, "P", None)
    if cached is not None:
        out["coarse_P"] = np.asarray(cached, dtype=np.float64)
    if extra:
        meta["extra"] = sorted(extra)
        for k, v in extra.items():
            out["X_" + k] = np.asarray(v)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    (np.savez_compressed if compressed else np.savez)(path, **out)


def load_hierarchy(path, device=0, stream=None):
    """Read a hierarchy written by ``save_hierarchy``. Returns ``(MultilevelSolver, extra_dict)``."""
    import ast
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    levels = []
    for k, m in enumerate(meta["levels"]):
        lvl = MultilevelSolver.Level()
        lvl.A = _get_matrix(z, f"L{k}_A", m["A"])
        if "P" in m:
            lvl.P = _get_matrix(z, f"L{k}_P", m["P"])
            lvl.R = _get_matrix(z, f"L{k}_R", m["R"])
            lvl.presmoother = _get_smoother(z, f"L{k}_pre", m["pre"], lvl)
            lvl.postsmoother = _get_smoother(z, f"L{k}_post", m["post"], lvl)
        levels.append(lvl)
    ml = MultilevelSolver(levels, coarse_solver=ast.literal_eval(meta["coarse_solver"]),
                          device=device, stream=stream)
    ml.symmetric_smoothing = meta.get("symmetric_smoothing", False)
    if "coarse_P" in z:
        ml.coarse_solver.P = np.ascontiguousarray(z["coarse_P"], dtype=np.float64)
    extra = {k: z["X_" + k] for k in meta.get("extra", [])}
    return ml, extra
