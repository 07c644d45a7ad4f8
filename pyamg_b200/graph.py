"""Graph helpers used at smoother-setup time (host): vertex colouring for multi-colour GS.

Plays the role of pyamg.graph.vertex_coloring (pyamg/graph.py:84-126); built in are first-fit greedy
colourings in natural, smallest-last (Matula-Beck; fewest colours on the dense coarse operators, the
smoother factory's default) and largest-degree-first order -- all red-black on 5/7-point stencils.  Any colouring --
including the reference's 'MIS' one -- gives a valid multi-colour sweep: the engine derives the
dependency waves from the row list itself (csrc/engine.cu build_waves), so a colouring only
influences speed, never correctness, and the CPU oracle sweeps the same row list sequentially.
"""
import numpy as np
from scipy import sparse

from . import _host as H


def vertex_coloring(G, method="greedy"):
    """Colours (int32 array, starting at 0) such that no edge of G joins equal colours."""
    orders = {"greedy": 0, "natural": 0, "smallest_last": 1, "SL": 1, "LDF": 2}
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

    colors, ok = run(G)
    if not ok:     # structurally non-symmetric operator: colour the symmetrised pattern instead
        colors, ok = run((abs(G) + abs(G).T).tocsr())
    return colors


def color_order(colors):
    """Row list sorted by colour (stable): the ``indices`` of gauss_seidel_indexed."""
    return np.argsort(colors, kind="stable").astype(np.int32)
