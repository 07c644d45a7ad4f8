"""Graph helpers used at smoother-setup time (host): vertex colouring for multi-colour GS.

Plays the role of pyamg.graph.vertex_coloring (pyamg/graph.py:84-126); built in are the reference's 'MIS' colouring
(= first-fit greedy in natural order, see below), first-fit in smallest-last (Matula-Beck; fewest colours on the dense coarse operators, the
smoother factory's default) and largest-degree-first order -- all red-black on 5/7-point stencils.  Any colouring --
including the reference's 'MIS' one -- gives a valid multi-colour sweep: the engine derives the
dependency waves from the row list itself (csrc/engine.cu build_waves), so a colouring only
influences speed, never correctness, and the CPU oracle sweeps the same row list sequentially.
"""
import numpy as np
from scipy import sparse

from . import _host as H


def _device_coloring_wanted(where):
    import os
    if where is not None:
        if where not in ("host", "gpu"):
            raise ValueError("where must be 'host' or 'gpu'")
        return where == "gpu"
    env = os.environ.get("AMGB_GPU_COLORING")
    if env in ("0", "1"):
        return env == "1"
    # default: the host routine.  Measured on a B200 box (profiles/r02_widening.jsonl): the multi-threaded host first
    # fit colours the 2.1 M-row / 14.6 M-entry level in 0.031 s, the device wavefront needs 0.033 s end to end (graph
    # upload included) and 0.30 s on the 1 M-row / 19.6 M-entry level with its longer dependency chains -- identical
    # colours, no gain: the colouring stays on the host unless asked for (where='gpu', AMGB_GPU_COLORING=1).
    return False


def _mis_coloring_device(G):
    """The 'MIS' colouring on the device (csrc/coloring.cuh: wavefront of first-fit decisions, colours identical to
    the sequential reference vertex by vertex).  Returns None when it needs more than 256 colours."""
    import ctypes
    from . import _engine as E
    n = G.shape[0]
    Ap = np.ascontiguousarray(G.indptr, dtype=np.int32)
    Aj = np.ascontiguousarray(G.indices, dtype=np.int32)
    colors = np.empty(n, dtype=np.int32)
    k, rounds = ctypes.c_int32(0), ctypes.c_int32(0)
    try:
        E.check(E.lib().amgb_host_vertex_coloring_mis(n, E.i32p(Ap), E.i32p(Aj), E.i32p(colors), ctypes.byref(k),
                                                      ctypes.byref(rounds)))
    except NotImplementedError:
        return None
    return colors


def vertex_coloring(G, method="greedy", where=None):
    """Colours (int32 array, starting at 0) such that no edge of G joins equal colours.

    ``where='gpu'`` (or AMGB_GPU_COLORING=1) computes the 'MIS' / natural-order greedy colouring on the device: same
    colours vertex by vertex; the host routine is the default because it is as fast (see _device_coloring_wanted)."""
    # 'MIS' (pyamg.graph.vertex_coloring's default, amg_core/graph.h:218-235): colour k is the lexicographically first
    # maximal independent set of what colours 0..k-1 left over (maximal_independent_set_serial walks the vertices in
    # index order, graph.h:128-199).  Vertex i misses set k exactly when a SMALLER-index neighbour is in it, so i gets
    # the smallest colour no smaller-index neighbour holds: natural-order first fit.  Same colours vertex by vertex
    # (tests/test_setup.py::test_mis_colouring_is_the_reference's, against the real reference).
    orders = {"greedy": 0, "natural": 0, "MIS": 0, "smallest_last": 1, "SL": 1, "LDF": 2}
    if method not in orders:
        raise NotImplementedError(f"colouring method {method!r}: built in are 'greedy' (natural-order first "
                                  "fit), 'smallest_last', 'LDF'; or pass the row list from "
                                  "pyamg.graph.vertex_coloring explicitly")
    G = sparse.csr_array(G)
    if G.shape[0] != G.shape[1]:
        raise ValueError("expected square matrix")
    n = G.shape[0]

    def run(M):
        Ap = np.ascontiguousarray(M.indptr, dtype=np.int32)
        Aj = np.ascontiguousarray(M.indices, dtype=np.int32)
        colors = np.empty(n, dtype=np.int32)
        H.lib().amgb_setup_greedy_coloring_ordered(n, H.ip(Ap), H.ip(Aj), orders[method], H.ip(colors))
        return colors, bool(H.lib().amgb_setup_coloring_is_valid(n, H.ip(Ap), H.ip(Aj), H.ip(colors)))

    if orders[method] == 0 and _device_coloring_wanted(where):
        Ap = np.ascontiguousarray(G.indptr, dtype=np.int32)
        Aj = np.ascontiguousarray(G.indices, dtype=np.int32)
        colors = _mis_coloring_device(G)
        if colors is not None and H.lib().amgb_setup_coloring_is_valid(n, H.ip(Ap), H.ip(Aj), H.ip(colors)):
            return colors          # (an invalid result means a structurally non-symmetric pattern: host path below)
    colors, ok = run(G)
    if not ok:     # structurally non-symmetric operator: colour the symmetrised pattern instead
        colors, ok = run((abs(G) + abs(G).T).tocsr())
    return colors


def color_order(colors):
    """Row list sorted by colour (stable): the ``indices`` of gauss_seidel_indexed."""
    return np.argsort(colors, kind="stable").astype(np.int32)
