"""Multi-GPU V-cycle: the big (fine) levels row-partitioned across ranks, the rest replicated.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  The reference has no
distributed path; this layer adds the decomposition the north star names:

* every level whose operator has more than ``dist_nnz`` stored entries is partitioned into contiguous
  row blocks (natural ordering of that level: gallery grids number the last dimension fastest, so a
  block is a slab and its halo one grid plane per side); all coarser levels are REPLICATED -- every
  rank runs the same cycle on them (no broadcast needed), "replicas only" below the partition;
* before every operator application that gathers a partitioned vector (Gauss-Seidel wave, Jacobi
  sweep, residual, restriction of a partitioned coarse level, prolongation from one) the ranks
  exchange the boundary entries of that vector with ONE all-gather (``ncclAllGather``): rank p packs
  the entries other ranks reference (``send_idx``), the gathered blocks land directly behind the
  owned part of the vector, and the local operator's column indices already point there -- no unpack;
* restriction onto a replicated level is a partial product ``R[:, owned] r_owned`` followed by
  ``ncclAllReduce(sum)``; the residual norm of the stop test is an 8-byte all-reduce.

Numerics: local rows keep the reference's per-row entry order, Gauss-Seidel waves are the GLOBAL
dependency waves, so the iterates equal the single-GPU ones except for the summation order of the
all-reduced restriction (1e-16 relative; parity bar 1e-12).

The arithmetic is delegated to a backend (``GpuBackend`` here: sm_100a tile kernels through
``amgb_operator_*`` + NCCL).  The partitioning / halo-plan code is pure NumPy and is exercised on
CPU with gloo (tests/test_dist_cpu.py) against the sequential oracle.
"""
import ctypes

import numpy as np
from scipy import sparse

from . import _engine as E
from .relaxation import smoothing

OP_SPMV, OP_RESID, OP_PADD, OP_JACOBI, OP_GS = 0, 1, 2, 3, 4


def block_bounds(n, world):
    """Contiguous, balanced row blocks: rank p owns [bounds[p], bounds[p+1])."""
    return np.array([(n * p) // world for p in range(world + 1)], dtype=np.int64)


def coarse_bounds_from_splitting(splitting, fine_bounds):
    """Coarse blocks aligned with the fine slabs: rank p owns the C points of its fine rows."""
    csum = np.concatenate([[0], np.cumsum(np.asarray(splitting, dtype=np.int64))])
    return csum[fine_bounds]


def wave_schedule(A, order=None):
    """1-based dependency-wave id of every position of the sequential sweep over `order` (engine's rule)."""
    A = sparse.csr_array(A)
    n = A.shape[0]
    Ap = np.ascontiguousarray(A.indptr, dtype=np.int32)
    Aj = np.ascontiguousarray(A.indices, dtype=np.int32)
    m = n if order is None else len(order)
    lst = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
    wave = np.empty(m, dtype=np.int32)
    nw = ctypes.c_int32(0)
    E.check(E.lib().amgb_wave_schedule(n, E.i32p(Ap), E.i32p(Aj), None if lst is None else E.i32p(lst), m,
                                       E.i32p(wave), ctypes.byref(nw)))
    return wave, int(nw.value)


class _Space:
    """A partitioned vector space (one hierarchy level): ownership, local order, halo plan."""

    def __init__(self, n, bounds, rank, mode="allgather"):
        self.mode = mode
        self.n, self.bounds, self.rank = n, np.asarray(bounds, dtype=np.int64), rank
        self.world = len(bounds) - 1
        self.lo, self.hi = int(bounds[rank]), int(bounds[rank + 1])
        self.n_own = self.hi - self.lo
        self.local_of_owned = np.arange(self.n_own, dtype=np.int64)   # owned global j -> local index (j - lo)
        self.needs = [set() for _ in range(self.world)]              # filled by add_reader
        self._remote = [[] for _ in range(self.world)]

    def add_reader(self, M, row_bounds):
        """Matrix M (rows partitioned by row_bounds) gathers vectors of this space: record, for every
        rank q, the columns it references but does not own."""
        M = sparse.csr_array(M)
        for q in range(self.world):
            r0, r1 = int(row_bounds[q]), int(row_bounds[q + 1])
            cols = M.indices[M.indptr[r0]:M.indptr[r1]]
            lo, hi = int(self.bounds[q]), int(self.bounds[q + 1])
            rem = cols[(cols < lo) | (cols >= hi)]
            if rem.size:
                self._remote[q].append(np.unique(rem))

    def finish(self):
        """Boundary sets B_p (what each rank must send), padded size, and the send list of this rank."""
        need = [np.unique(np.concatenate(r)) if r else np.empty(0, dtype=np.int64) for r in self._remote]
        allneed = np.unique(np.concatenate(need)) if any(x.size for x in need) else np.empty(0, dtype=np.int64)
        self.B = [allneed[(allneed >= self.bounds[p]) & (allneed < self.bounds[p + 1])].astype(np.int64)
                  for p in range(self.world)]
        self.maxB = max(1, max(len(b) for b in self.B))
        # neighbour ("p2p") plan: what THIS rank receives from every owner p, and what it sends to every q
        r = self.rank
        self.recv_from = [need[r][(need[r] >= self.bounds[p]) & (need[r] < self.bounds[p + 1])].astype(np.int64)
                          if p != r else np.empty(0, dtype=np.int64) for p in range(self.world)]
        self.send_to = [need[q][(need[q] >= self.lo) & (need[q] < self.hi)].astype(np.int64)
                        if q != r else np.empty(0, dtype=np.int64) for q in range(self.world)]
        self.recv_off = np.concatenate([[0], np.cumsum([len(a) for a in self.recv_from])]).astype(np.int64)
        self.send_off = np.concatenate([[0], np.cumsum([len(a) for a in self.send_to])]).astype(np.int64)
        if self.mode == "p2p":
            self.n_ext = self.n_own + max(1, int(self.recv_off[-1]))
        else:
            self.n_ext = self.n_own + self.world * self.maxB
        del self._remote

    def set_local_order(self, perm_global_rows):
        """perm_global_rows: owned GLOBAL indices in the order they are stored locally."""
        pos = np.empty(self.n_own, dtype=np.int64)
        pos[np.asarray(perm_global_rows, dtype=np.int64) - self.lo] = np.arange(self.n_own)
        self.local_of_owned = pos
        self.order = np.asarray(perm_global_rows, dtype=np.int64)

    def send_idx(self):
        if self.mode == "p2p":      # packed per destination rank, ascending
            ids = np.concatenate(self.send_to) if self.send_off[-1] else np.empty(0, dtype=np.int64)
            return self.local_of_owned[ids - self.lo].astype(np.int32)
        return self.local_of_owned[self.B[self.rank] - self.lo].astype(np.int32)

    def map_cols(self, cols):
        """global column ids -> indices into this rank's extended vector [owned | gathered blocks]."""
        cols = np.asarray(cols, dtype=np.int64)
        out = np.empty(cols.shape, dtype=np.int64)
        own = (cols >= self.lo) & (cols < self.hi)
        out[own] = self.local_of_owned[cols[own] - self.lo]
        rem = ~own
        if rem.any():
            rc = cols[rem]
            q = np.searchsorted(self.bounds, rc, side="right") - 1
            off = np.empty(rc.shape, dtype=np.int64)
            for p in np.unique(q):
                m = q == p
                table = self.recv_from[p] if self.mode == "p2p" else self.B[p]
                k = np.searchsorted(table, rc[m])
                if np.any(k >= len(table)) or np.any(table[np.minimum(k, len(table) - 1)] != rc[m]):
                    raise RuntimeError("halo plan is missing a referenced column")
                base = self.recv_off[p] if self.mode == "p2p" else p * self.maxB
                off[m] = self.n_own + base + k
            out[rem] = off
        return out


class _Replicated:
    """A replicated vector space (every rank holds all n entries)."""

    def __init__(self, n):
        self.n = self.n_own = self.n_ext = n
        self.world = 1

    def map_cols(self, cols):
        return np.asarray(cols, dtype=np.int64)


def _local_rows(M, rows_global, colspace, n_cols_ext):
    """Rows `rows_global` (in that order) of M with columns renumbered into colspace's extended layout."""
    M = sparse.csr_array(M)
    rows_global = np.asarray(rows_global, dtype=np.int64)
    starts, ends = M.indptr[rows_global], M.indptr[rows_global + 1]
    lens = (ends - starts).astype(np.int64)
    indptr = np.zeros(len(rows_global) + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    take = np.repeat(starts.astype(np.int64) - indptr[:-1], lens) + np.arange(indptr[-1], dtype=np.int64)
    cols = colspace.map_cols(M.indices[take])
    return sparse.csr_array((np.ascontiguousarray(M.data[take], dtype=np.float64), cols.astype(np.int32),
                             indptr.astype(np.int32)), shape=(len(rows_global), n_cols_ext))


class DistLevel:
    pass


def build_plan(ml, world, rank, n_dist=None, dist_nnz=20_000_000, halo="allgather"):
    """Partition the leading levels of hierarchy `ml` for (world, rank). Returns (levels, n_dist).

    Pure host code; every rank runs it on the same (replicated) host hierarchy and keeps its part."""
    nl = len(ml.levels)
    if n_dist is None:
        n_dist = 0
        while n_dist < nl - 1 and ml.levels[n_dist].A.nnz > dist_nnz:
            n_dist += 1
        n_dist = max(n_dist, 1) if nl > 1 else 0
    n_dist = min(n_dist, nl - 1)
    # ownership: level 0 in balanced slabs; a coarse level follows its parent's slabs when a C/F
    # splitting is known (classical AMG), else balanced blocks
    bounds = [block_bounds(ml.levels[0].A.shape[0], world)]
    for l in range(1, n_dist):
        sp = getattr(ml.levels[l - 1], "splitting", None)
        if sp is not None and len(sp) == ml.levels[l - 1].A.shape[0] and int(np.sum(sp)) == ml.levels[l].A.shape[0]:
            bounds.append(coarse_bounds_from_splitting(sp, bounds[l - 1]))
        else:
            bounds.append(block_bounds(ml.levels[l].A.shape[0], world))
    if halo not in ("allgather", "p2p", "peer"):
        raise ValueError("halo must be 'allgather', 'p2p' or 'peer'")
    # 'peer' (stores into the neighbours' memory over NVLink, amgb_comm_*) shares the neighbour layout of 'p2p'
    spaces = [_Space(ml.levels[l].A.shape[0], bounds[l], rank, mode="p2p" if halo == "peer" else halo)
              for l in range(n_dist)]
    for l in range(n_dist):
        lvl = ml.levels[l]
        spaces[l].add_reader(lvl.A, bounds[l])                         # smoother / residual gather x_l
        if l + 1 < n_dist:
            spaces[l].add_reader(lvl.R, bounds[l + 1])                 # restriction gathers r_l
            spaces[l + 1].add_reader(lvl.P, bounds[l])                 # prolongation gathers x_{l+1}
    for s in spaces:
        s.finish()

    out = []
    for l in range(n_dist):
        lvl, sp = ml.levels[l], spaces[l]
        D = DistLevel()
        D.space = sp
        keep = []
        D.pre = smoothing.describe(getattr(lvl, "presmoother", None), lvl.A, keep)
        D.post = smoothing.describe(getattr(lvl, "postsmoother", None), lvl.A, keep)
        D._keep = keep
        for S in (D.pre, D.post):
            if S.kind not in (E.SM_NONE, E.SM_JACOBI, E.SM_GAUSS_SEIDEL, E.SM_POLYNOMIAL):
                raise NotImplementedError("partitioned levels run Jacobi, Gauss-Seidel (global waves) and polynomial "
                                          f"smoothers; smoother kind {S.kind} needs the single-GPU engine")
            if S.kind == E.SM_POLYNOMIAL:      # the coefficient array must outlive the descriptor
                S._coef = np.ctypeslib.as_array(S.coefficients, shape=(S.n_coefficients,)).copy()
        # global dependency waves of the (pre) Gauss-Seidel sweep decide the local storage order
        D.wave_ptr = None
        gs = D.pre if D.pre.kind == E.SM_GAUSS_SEIDEL else (D.post if D.post.kind == E.SM_GAUSS_SEIDEL else None)
        owned = np.arange(sp.lo, sp.hi, dtype=np.int64)
        if gs is not None:
            n = lvl.A.shape[0]
            lst = None
            if gs.n_indices:
                lst = np.ctypeslib.as_array(gs.indices, shape=(gs.n_indices,)).copy()
                if len(lst) != n or len(np.unique(lst)) != n:
                    raise NotImplementedError("partitioned Gauss-Seidel needs a row list that is a permutation")
            for other in (D.pre, D.post):
                if other.kind == E.SM_GAUSS_SEIDEL and other is not gs:
                    lo = None if not other.n_indices else np.ctypeslib.as_array(other.indices, shape=(other.n_indices,))
                    same = (lo is None and lst is None) or (lo is not None and lst is not None and np.array_equal(lo, lst))
                    if not same:
                        raise NotImplementedError("pre/post Gauss-Seidel with different row lists on a partitioned level")
            wave_pos, nw = wave_schedule(lvl.A, lst)
            wave_of_row = np.empty(n, dtype=np.int32)
            wave_of_row[np.arange(n) if lst is None else lst] = wave_pos
            w_own = wave_of_row[sp.lo:sp.hi]
            order = np.argsort(w_own, kind="stable")
            owned = owned[order]
            D.wave_ptr = np.concatenate([[0], np.cumsum(np.bincount(w_own - 1, minlength=nw))]).astype(np.int64)
            D.n_waves = nw
        sp.set_local_order(owned)
        out.append(D)
    for l in range(n_dist):
        lvl, sp, D = ml.levels[l], spaces[l], out[l]
        D.A = _local_rows(lvl.A, sp.order, sp, sp.n_ext)
        D.send_idx = sp.send_idx()
        if l + 1 < n_dist:
            nxt = spaces[l + 1]
            D.next_partitioned = True
            D.P = _local_rows(lvl.P, sp.order, nxt, nxt.n_ext)
            D.R = _local_rows(lvl.R, nxt.order, sp, sp.n_ext)
        else:
            D.next_partitioned = False
            nc = lvl.P.shape[1]
            D.P = _local_rows(lvl.P, sp.order, _Replicated(nc), nc)
            # partial restriction: all coarse rows, owned fine columns only (local numbering)
            Rcsc = sparse.csr_array(lvl.R)[:, sp.lo:sp.hi].tocsr()
            Rcsc.indices = sp.local_of_owned[Rcsc.indices].astype(np.int32)
            D.R = sparse.csr_array((Rcsc.data, Rcsc.indices, Rcsc.indptr), shape=(nc, sp.n_own))
    return out, n_dist


# ------------------------------------------------------------------------------------------------
# cycle driver (backend-agnostic)
# ------------------------------------------------------------------------------------------------
class DistributedSolver:
    """V-cycles with the leading levels partitioned across ``backend.world`` ranks.

    ``solve_device``-style use: ``load(b_global)``, ``cycles(k)``, ``gather_x()``; residual norms are
    all-reduced.  Only V-cycles (the partitioned recursion has one coarse visit per level)."""

    def __init__(self, ml, backend, n_dist=None, dist_nnz=20_000_000, halo="allgather"):
        """halo='allgather' (default; one ncclAllGather of the padded boundary blocks per exchange) or 'p2p'
        (grouped send/recv with exactly the ranks that reference each other's entries: moves
        world/#neighbours times fewer bytes; host logic covered by the gloo tests)."""
        from .multilevel import MultilevelSolver
        self.ml, self.be, self.halo_mode = ml, backend, halo
        self._graph = None
        self.plan, self.n_dist = build_plan(ml, backend.world, backend.rank, n_dist=n_dist, dist_nnz=dist_nnz,
                                            halo=halo)
        be = backend
        self.lv = []
        for D in self.plan:
            sp = D.space
            L = DistLevel()
            L.D, L.sp = D, sp
            L.A = be.operator(D.A, D.wave_ptr)
            L.P = be.operator(D.P, None)
            L.R = be.operator(D.R, None)
            L.send_idx = be.index(D.send_idx)
            L.send = be.vector(max(sp.maxB, int(sp.send_off[-1])))
            L.x, L.xalt, L.b, L.r = (be.vector(sp.n_ext) for _ in range(4))
            L.x_home = L.x
            L.poly = be.vector(sp.n_ext) if E.SM_POLYNOMIAL in (D.pre.kind, D.post.kind) else None
            self.lv.append(L)
        if halo == "peer" and be.world > 1:
            # one communicator for all levels: neighbour set = union over the levels (symmetric by construction: q
            # sends to p exactly when p receives from q), staging sized for the largest halo
            mask, cap = 0, 1
            for L in self.lv:
                sp = L.sp
                cap = max(cap, int(sp.recv_off[-1]))
                for q in range(be.world):
                    if sp.send_off[q + 1] > sp.send_off[q] or sp.recv_off[q + 1] > sp.recv_off[q]:
                        mask |= 1 << q
            recv_offs = be.allgather_object([np.asarray(L.sp.recv_off, dtype=np.int64) for L in self.lv])
            for l, L in enumerate(self.lv):
                # where MY block starts inside rank q's halo region of level l
                L.peer_off = np.array([int(recv_offs[q][l][be.rank]) for q in range(be.world)], dtype=np.int64)
                L.send_off = np.ascontiguousarray(L.sp.send_off, dtype=np.int64)
            cap = max(be.allgather_object(int(cap)))           # ONE staging size: peers address each other's buffers
            be.comm_setup(cap, mask)
        # replicated remainder: an ordinary engine hierarchy on every rank
        self.sub = backend.sub_solver(MultilevelSolver, ml, self.n_dist)
        nrep = ml.levels[self.n_dist].A.shape[0]
        self.bc_rep, self.xc_rep = be.vector(nrep), be.vector(nrep)
        self.norm2 = be.vector(1)
        self.launches = 0

    # -- communication --------------------------------------------------------------------------
    def halo(self, L, v):
        """All-gather the boundary entries of partitioned vector v into its halo region."""
        if self.be.world == 1:
            return
        self.n_exchanges = getattr(self, "n_exchanges", 0) + 1
        if self.halo_mode == "peer":      # ONE kernel: pack + stores into the neighbours' staging + flags + unpack
            self.be.exchange_peer(v, L.sp.n_own, L.send_idx, L.send_off, L.peer_off, int(L.sp.recv_off[-1]))
            return
        self.be.gather(v, L.send_idx, L.send, len(L.D.send_idx))
        if self.halo_mode == "p2p":
            self.be.exchange(L.send, L.sp.send_off, v, L.sp.n_own, L.sp.recv_off)
        else:
            self.be.allgather(L.send, v, L.sp.n_own, L.sp.maxB)

    # -- smoothers --------------------------------------------------------------------------------
    def smooth(self, L, S):
        be = self.be
        if S.kind == E.SM_NONE:
            return
        if S.kind == E.SM_JACOBI:
            for _ in range(S.iterations):
                self.halo(L, L.x)
                be.apply(L.A, OP_JACOBI, L.x, L.b, L.xalt, omega=S.omega)
                L.x, L.xalt = L.xalt, L.x
            return
        if S.kind == E.SM_POLYNOMIAL:
            # relaxation.polynomial (relaxation.py:646-659): r = b - A x; h = c_0 r; h = c_k r + A h; x += h.
            # Every SpMV of a partitioned vector is preceded by its halo exchange.
            coef = S._coef
            for _ in range(S.iterations):
                self.halo(L, L.x)
                be.apply(L.A, OP_RESID, L.x, L.b, L.r)
                h, Ah = L.xalt, L.poly
                be.scale_to(h, float(coef[0]), L.r)
                for c in coef[1:]:
                    self.halo(L, h)
                    be.apply(L.A, OP_SPMV, h, None, Ah)
                    be.axpby(float(c), L.r, 1.0, Ah)
                    h, Ah = Ah, h
                be.axpby(1.0, h, 1.0, L.x)
            return
        if S.kind != E.SM_GAUSS_SEIDEL:
            raise NotImplementedError(f"smoother kind {S.kind} on a partitioned level")
        nw = L.D.n_waves
        om = 1.0 if S.sweep == E.SWEEPS["symmetric"] else S.omega
        for _ in range(S.iterations):
            seq = []
            if S.sweep in (E.SWEEPS["forward"], E.SWEEPS["symmetric"]):
                seq += list(range(nw))
            if S.sweep == E.SWEEPS["backward"]:
                seq += list(range(nw - 1, -1, -1))
            if S.sweep == E.SWEEPS["symmetric"]:
                seq += list(range(nw - 2, -1, -1))      # the repeated middle wave is idempotent
            for w in seq:
                self.halo(L, L.x)
                be.apply(L.A, OP_GS, L.x, L.b, L.x, omega=om, wave=w)

    # -- one V-cycle on partitioned level l (multilevel.py:584-662) --------------------------------
    def cycle(self, l):
        be, L = self.be, self.lv[l]
        D = L.D
        self.smooth(L, D.pre)
        self.halo(L, L.x)
        be.apply(L.A, OP_RESID, L.x, L.b, L.r)
        if D.next_partitioned:
            C = self.lv[l + 1]
            self.halo(L, L.r)
            be.apply(L.R, OP_SPMV, L.r, None, C.b)
            be.fill(C.x, 0.0)
            self.cycle(l + 1)
            self.halo(C, C.x)
            be.apply(L.P, OP_PADD, C.x, None, L.x)
        else:
            be.apply(L.R, OP_SPMV, L.r, None, self.bc_rep)
            be.allreduce(self.bc_rep)
            be.fill(self.xc_rep, 0.0)
            self.sub.cycle_device(self.bc_rep, self.xc_rep)
            be.apply(L.P, OP_PADD, self.xc_rep, None, L.x)
        self.smooth(L, D.post)
        if L.x is not L.x_home:        # odd number of Jacobi ping-pongs: bring the iterate home (a replayed graph
            self.be.copy_owned(L.x_home, L.x, L.sp.n_own)          # has its buffer addresses baked in)
            L.x, L.xalt = L.x_home, L.x

    # -- public -------------------------------------------------------------------------------------
    def load(self, b_global, x0_global=None):
        L = self.lv[0]
        sp = L.sp
        self.be.set_owned(L.b, np.asarray(b_global, dtype=np.float64)[sp.order])
        if x0_global is None:
            self.be.fill(L.x, 0.0)
        else:
            self.be.set_owned(L.x, np.asarray(x0_global, dtype=np.float64)[sp.order])

    def residual_norm(self):
        L = self.lv[0]
        self.halo(L, L.x)
        self.be.apply(L.A, OP_RESID, L.x, L.b, L.r, norm2=self.norm2)
        self.be.allreduce(self.norm2)
        return self.norm2

    def cycles(self, k, norms=None):
        """k V-cycles; if `norms` (backend vector, k+1 slots) is given the residual norms^2 are stored."""
        if norms is not None:
            self.be.copy_scalar(self.residual_norm(), norms, 0)
        for it in range(1, k + 1):
            if self.n_dist == 0:
                raise RuntimeError("nothing is partitioned")
            if self._graph is not None:
                self._graph.replay()
            else:
                self.cycle(0)
            if norms is not None:
                self.be.copy_scalar(self.residual_norm(), norms, it)

    def capture_graph(self):
        """Capture one distributed V-cycle -- engine kernels, peer-memory halo exchanges (amgb_comm_exchange), the
        all-reduce of the restriction and the replicated sub-hierarchy's own graph -- into a CUDA graph on the
        backend's stream, so the ~500 host-issued launches per cycle become one replay.  Requires warmed-up
        cycles (all engine graphs instantiated)."""
        torch = getattr(self.be, "torch", None)
        if torch is None:
            raise NotImplementedError("graph capture needs the GPU backend")
        x0 = getattr(self, "n_exchanges", 0)
        self.cycle(0)                                  # make sure every lazily built piece exists
        self.exchanges_per_cycle = getattr(self, "n_exchanges", 0) - x0
        torch.cuda.synchronize()
        l0 = self.be.kernel_launches
        g = torch.cuda.CUDAGraph()
        # thread-local capture mode: NCCL's watchdog thread keeps querying its events while this thread captures
        with torch.cuda.graph(g, stream=self.be.stream, capture_error_mode="thread_local"):
            self.cycle(0)
        self.graph_launches = self.be.kernel_launches - l0      # kernels one replay launches
        self._graph = g
        return g

    def gather_x(self):
        """The full iterate on every rank (host array, original numbering)."""
        L = self.lv[0]
        sp = L.sp
        own = self.be.get_owned(L.x, sp.n_own)
        nat = np.empty(sp.n_own, dtype=np.float64)
        nat[sp.order - sp.lo] = own                      # back to the natural order of the owned block
        return self.be.allgather_host(nat)


# ------------------------------------------------------------------------------------------------
# GPU backend: engine tile kernels + NCCL
# ------------------------------------------------------------------------------------------------
class GpuBackend:
    """torch CUDA tensors for memory, ``amgb_operator_*`` kernels for arithmetic, torch.distributed
    (NCCL) for the collectives.  world == 1 works without an initialised process group."""

    def __init__(self, device=0, rank=0, world=1, group=None):
        import torch
        self.torch = torch
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device("cuda", device)
        self.dev_index = device
        E.require_gpu()
        self.L = E.lib()
        # every torch op, NCCL collective and engine kernel of this backend must share ONE stream; the
        # engine treats a NULL handle as "create my own", so never hand it torch's default stream
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.current_stream(self.device)
        if self.stream.cuda_stream == 0:
            self.stream = torch.cuda.Stream(device=self.device)
            torch.cuda.set_stream(self.stream)
        self._ops = []
        self.kernel_launches = 0

    # memory
    def vector(self, n):
        return self.torch.zeros(int(n) + 2, dtype=self.torch.float64, device=self.device)   # +2: TMA padding

    def index(self, idx):
        return self.torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int32)).to(self.device)

    def fill(self, v, val):
        v.fill_(val)

    def set_owned(self, v, host):
        v[:len(host)].copy_(self.torch.from_numpy(np.ascontiguousarray(host)))

    def get_owned(self, v, n):
        return v[:n].cpu().numpy()

    def copy_scalar(self, src, dst, slot):
        dst[slot:slot + 1].copy_(src[:1])

    def copy_owned(self, dst, src, n):
        dst[:n].copy_(src[:n])

    def scale_to(self, dst, a, src):                # dst = a * src (owned part and halo region alike)
        self.torch.mul(src, a, out=dst)

    def axpby(self, a, x, b, y):                    # y = a x + b y
        y.mul_(b).add_(x, alpha=a)

    # operators
    def operator(self, M, wave_ptr):
        keep = []
        Mc = E.as_matrix(M, keep)
        h = ctypes.c_void_p()
        wp, nw = None, 0
        if wave_ptr is not None:
            wp = np.ascontiguousarray(wave_ptr, dtype=np.int64)
            nw = len(wp) - 1
        E.check(self.L.amgb_operator_create(self.dev_index, ctypes.byref(Mc),
                                            None if wp is None else wp.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                            nw, ctypes.c_void_p(self.stream.cuda_stream), ctypes.byref(h)))
        self._ops.append(h)
        return h

    def apply(self, op, kind, x, b, y, r=None, omega=0.0, wave=-1, norm2=None):
        P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        E.check(self.L.amgb_operator_apply(op, kind, P(x), P(b), P(y), P(r), float(omega), P(norm2), int(wave)))
        self.kernel_launches += 1

    def gather(self, v, idx, out, n):
        if n > 0:
            E.check(self.L.amgb_dev_gather(ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(idx.data_ptr()),
                                           ctypes.c_void_p(out.data_ptr()), int(n),
                                           ctypes.c_void_p(self.stream.cuda_stream)))
            self.kernel_launches += 1

    # collectives
    def allgather(self, send, v, n_own, maxB):
        import torch.distributed as dist
        dist.all_gather_into_tensor(v[n_own:n_own + self.world * maxB], send[:maxB], group=self.group)

    def exchange(self, send, send_off, v, n_own, recv_off):
        """Grouped NCCL send/recv with the ranks that share boundary entries (halo='p2p')."""
        import torch.distributed as dist
        ops = []
        for q in range(self.world):
            ns = int(send_off[q + 1] - send_off[q])
            nr = int(recv_off[q + 1] - recv_off[q])
            if q == self.rank:
                continue
            if ns:
                ops.append(dist.P2POp(dist.isend, send[int(send_off[q]):int(send_off[q + 1])], q, group=self.group))
            if nr:
                ops.append(dist.P2POp(dist.irecv, v[n_own + int(recv_off[q]):n_own + int(recv_off[q + 1])], q,
                                      group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def allreduce(self, v):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(v[:v.numel() - 2], group=self.group)

    def allgather_object(self, obj):
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, obj, group=self.group)
        return parts

    def comm_setup(self, cap, nbr_mask):
        """Peer-memory communicator (csrc/abi_comm.cuh): every rank exports one IPC block, the 64-byte handles
        travel through the process group once, neighbours map each other's blocks."""
        import torch.distributed as dist
        handle = ctypes.create_string_buffer(64)
        c = ctypes.c_void_p()
        E.check(self.L.amgb_comm_create(self.dev_index, self.world, self.rank, int(cap),
                                        ctypes.c_void_p(self.stream.cuda_stream), ctypes.byref(c), handle))
        self._comm = c
        handles = self.allgather_object(bytes(handle.raw))
        E.check(self.L.amgb_comm_connect(c, b"".join(handles), int(nbr_mask)))
        dist.barrier(group=self.group)

    def exchange_peer(self, v, n_own, send_idx, send_off, peer_off, recv_total):
        i64p = ctypes.POINTER(ctypes.c_int64)
        E.check(self.L.amgb_comm_exchange(self._comm, ctypes.c_void_p(v.data_ptr()), int(n_own),
                                          ctypes.c_void_p(send_idx.data_ptr()), send_off.ctypes.data_as(i64p),
                                          peer_off.ctypes.data_as(i64p), int(recv_total)))
        self.kernel_launches += 1

    def allgather_host(self, own):
        """Concatenation of every rank's owned block (host array) -- result path, not timed."""
        if self.world == 1:
            return own
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, own, group=self.group)
        return np.concatenate(parts)

    def sub_solver(self, MultilevelSolver, ml, first):
        sub = MultilevelSolver(list(ml.levels[first:]), coarse_solver="pinv", device=self.dev_index,
                               stream=self.stream.cuda_stream)
        sub.coarse_solver = ml.coarse_solver
        be = self

        class _Sub:
            def cycle_device(self_inner, b, x):
                E.check(be.L.amgb_solve_device(sub.handle, ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(x.data_ptr()),
                                               1, 0, 1, None))
                be.kernel_launches += sub.last_launches()
        self._sub = sub
        return _Sub()

    def close(self):
        for h in self._ops:
            self.L.amgb_operator_destroy(h)
        self._ops = []
        if getattr(self, "_comm", None) is not None:
            self.L.amgb_comm_destroy(self._comm)
            self._comm = None
