"""CPU model of the x-gather traffic of a colour wave: how many 32-byte sectors a warp-wide gather instruction
touches under the tile kernel's lane mapping (G lanes per row, 32/G rows per pass), per instruction, per tile
(~12 rows) and per CTA (~96 rows).  Used to compare colourings / in-wave row orders without a GPU:

    python tools/sector_model.py 128 1        # level 1 of the 128^3 RS hierarchy, greedy vs smallest-last colours
"""
import sys, numpy as np, scipy.sparse as sp
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def permuted(A, order):
    n=A.shape[0]; pos=np.empty(n,dtype=np.int64); pos[order]=np.arange(n)
    lens=np.diff(A.indptr)[order]; Bp=np.concatenate([[0],np.cumsum(lens)]).astype(np.int64)
    ent=np.repeat(A.indptr[order].astype(np.int64)-Bp[:-1],lens)+np.arange(Bp[-1])
    return Bp, pos[A.indices[ent]]
def model(Bp,Bj,G,rows,name, wave_of_pos=None):
    r0,r1=rows; jb,je=Bp[r0],Bp[r1]; cols=Bj[jb:je]; lens=np.diff(Bp[r0:r1+1])
    rid=np.repeat(np.arange(r1-r0),lens); k=np.arange(je-jb)-np.repeat(Bp[r0:r1]-jb,lens)
    rpp=32//G; inst=(rid//rpp).astype(np.int64)*4096+(k//G)
    ninst=len(np.unique(inst)); us=len(np.unique(inst*(1<<34)+cols//4))
    ut=len(np.unique((rid//12).astype(np.int64)*(1<<34)+cols//4)); uc=len(np.unique((rid//96).astype(np.int64)*(1<<34)+cols//4))
    print(f"{name} G={G}: sectors/inst {us/ninst:.2f}  secB/entry: inst {us*32/(je-jb):.1f} tile {ut*32/(je-jb):.1f} cta {uc*32/(je-jb):.1f}")
    return us*32/(je-jb)
def colour_order(A, method):
    from pyamg_b200.graph import vertex_coloring
    c=vertex_coloring(A, method=method)
    return c, np.argsort(c,kind='stable')


if __name__ == "__main__":
    import bench
    g, lev = int(sys.argv[1]), int(sys.argv[2])
    ml = bench.build_hierarchy((g, g, g))
    A = ml.levels[lev].A.tocsr()
    A.sort_indices()
    n = A.shape[0]
    Bp, Bj = permuted(A, np.arange(n))
    model(Bp, Bj, 1, (n // 2, min(n, n // 2 + 300000)), "natural order (no colouring)")
    for method in ("greedy", "smallest_last"):
        c, order = colour_order(A, method)
        cnt = np.bincount(c)
        starts = np.concatenate([[0], np.cumsum(cnt)])
        Bp, Bj = permuted(A, order)
        print(method, "colours", len(cnt))
        for w in (0, len(cnt) // 2):
            model(Bp, Bj, 2, (starts[w], starts[w + 1]), f"  wave {w}")
