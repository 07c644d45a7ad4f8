"""Worker of tests/test_dist_cpu.py: runs the distributed V-cycle of pyamg_b200.dist on CPU with the
gloo backend (world_size = WORLD_SIZE), arithmetic by a NumPy/oracle test backend, and writes the
gathered iterate + residual history of rank 0 to OUT.  Test infrastructure (may import oracle)."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle                                              # noqa: E402
from pyamg_b200 import dist as D                            # noqa: E402
from pyamg_b200.hierarchy_io import load_hierarchy          # noqa: E402


class NumpyBackend:
    """Same interface as pyamg_b200.dist.GpuBackend; vectors are NumPy arrays, collectives go through gloo."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def vector(self, n):
        return np.zeros(int(n) + 2)

    def index(self, idx):
        return np.asarray(idx, dtype=np.int64)

    def fill(self, v, val):
        v[:] = val

    def set_owned(self, v, host):
        v[:len(host)] = host

    def get_owned(self, v, n):
        return v[:n].copy()

    def copy_scalar(self, src, dst, slot):
        dst[slot] = src[0]

    def copy_owned(self, dst, src, n):
        dst[:n] = src[:n]

    def scale_to(self, dst, a, src):
        dst[:] = a * src

    def axpby(self, a, x, b, y):
        y *= b
        y += a * x

    def operator(self, M, wave_ptr):
        return (sp.csr_array(M), None if wave_ptr is None else np.asarray(wave_ptr))

    def apply(self, op, kind, x, b, y, r=None, omega=0.0, wave=-1, norm2=None):
        M, wp = op
        n, nc = M.shape
        if kind == D.OP_SPMV:
            y[:n] = M @ x[:nc]
        elif kind == D.OP_RESID:
            y[:n] = b[:n] - M @ x[:nc]
            if norm2 is not None:
                norm2[0] = float(np.dot(y[:n], y[:n]))
        elif kind == D.OP_PADD:
            y[:n] += M @ x[:nc]
        elif kind == D.OP_JACOBI:
            d = M.diagonal()                                  # local rows: the diagonal sits at local column i
            Ax = M @ x[:nc]
            xn = x[:n].copy()
            nz = d != 0
            xn[nz] = (1 - omega) * x[:n][nz] + omega * ((b[:n][nz] - (Ax[nz] - d[nz] * x[:n][nz])) / d[nz])
            y[:n] = xn
        else:                                                 # one Gauss-Seidel wave: rows are independent
            rows = np.arange(wp[wave], wp[wave + 1])
            if len(rows) == 0:
                return
            Mr = M[rows]
            d = M.diagonal()[rows]
            off = Mr @ x[:nc] - d * x[rows]
            g = (b[rows] - off) / np.where(d != 0, d, 1.0)
            new = g if omega == 1.0 else omega * g + (1 - omega) * x[rows]
            y[rows] = np.where(d != 0, new, x[rows])

    def gather(self, v, idx, out, n):
        out[:n] = v[idx[:n]]

    def allgather(self, send, v, n_own, maxB):
        parts = [torch.zeros(maxB, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(parts, torch.from_numpy(send[:maxB].copy()))
        v[n_own:n_own + self.world * maxB] = torch.cat(parts).numpy()

    def exchange(self, send, send_off, v, n_own, recv_off):
        reqs, bufs = [], []
        for q in range(self.world):
            if q == self.rank:
                continue
            ns, nr = int(send_off[q + 1] - send_off[q]), int(recv_off[q + 1] - recv_off[q])
            if ns:
                reqs.append(dist.isend(torch.from_numpy(send[int(send_off[q]):int(send_off[q + 1])].copy()), q))
            if nr:
                t = torch.zeros(nr, dtype=torch.float64)
                bufs.append((t, n_own + int(recv_off[q])))
                reqs.append(dist.irecv(t, q))
        for r in reqs:
            r.wait()
        for t, off in bufs:
            v[off:off + len(t)] = t.numpy()

    def allreduce(self, v):
        if self.world > 1:
            t = torch.from_numpy(v[:len(v) - 2].copy())
            dist.all_reduce(t)
            v[:len(v) - 2] = t.numpy()

    def allgather_host(self, own):
        parts = [None] * self.world
        dist.all_gather_object(parts, own)
        return np.concatenate(parts)

    def sub_solver(self, MultilevelSolver, ml, first):
        spec = oracle.hierarchy_spec(ml)[first:]
        cyc = oracle.Cycle(spec, coarse_pinv=ml.coarse_solver.P)

        class _Sub:
            def cycle_device(self_inner, b, x):
                n = spec[0]["A"].shape[0]
                x[:n] = cyc.solve(b[:n].copy(), tol=0, maxiter=1)
        return _Sub()


def main():
    name, n_dist, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    halo = sys.argv[4] if len(sys.argv) > 4 else "allgather"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ml, ex = load_hierarchy(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    be = NumpyBackend(rank, world)
    ds = D.DistributedSolver(ml, be, n_dist=n_dist, halo=halo)
    ds.load(ex["b"])
    ncyc = 4
    norms = np.zeros(ncyc + 1)
    ds.cycles(ncyc, norms=norms)
    x = ds.gather_x()
    info = {"maxB": [int(L.sp.maxB) for L in ds.lv], "n_own": [int(L.sp.n_own) for L in ds.lv]}
    if rank == 0:
        np.savez(out, x=x, res=np.sqrt(norms), maxB=np.array(info["maxB"]), n_own=np.array(info["n_own"]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
