"""Worker of tests/test_emulated_engine.py::test_distributed_cycle_on_the_emulator: the REAL multi-GPU layer
(pyamg_b200.dist.DistributedSolver with its GpuBackend code paths: amgb_operator_* tile kernels, packed halo
all-gather / neighbour send-recv, all-reduced restriction, the replicated sub-hierarchy through amgb_solve_device)
at world_size > 1 -- with torch CPU tensors as "device" memory, gloo as the collective library and the kernel
emulator (tests/emu) as the engine.  Test infrastructure."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
os.environ["AMGB_TEST_EMU"] = "1"
import conftest  # noqa: E402,F401  (installs the emulation library as the engine)
from pyamg_b200 import _engine as E  # noqa: E402
from pyamg_b200 import dist as D  # noqa: E402
from pyamg_b200.hierarchy_io import load_hierarchy  # noqa: E402


class _NullStream:
    cuda_stream = 0          # NULL handle: every engine object creates its own (emulated) stream


class EmuBackend(D.GpuBackend):
    """GpuBackend with the CUDA-specific plumbing replaced: CPU tensors, no stream, gloo collectives."""

    def __init__(self, rank, world):
        self.torch = torch
        self.rank, self.world, self.group = rank, world, None
        self.device = torch.device("cpu")
        self.dev_index = 0
        E.require_gpu()
        self.L = E.lib()
        self.stream = _NullStream()
        self._ops = []
        self._vectors = []
        self.kernel_launches = 0

    def vector(self, n):
        v = torch.zeros(int(n) + 2, dtype=torch.float64)
        self.L.amgb_emu_register_allocation(ctypes.c_void_p(v.data_ptr()), ctypes.c_size_t(v.numel() * 8))
        self._vectors.append(v)
        return v

    def close(self):
        for v in self._vectors:
            self.L.amgb_emu_unregister_allocation(ctypes.c_void_p(v.data_ptr()))
        super().close()

    def allgather(self, send, v, n_own, maxB):           # gloo has no all_gather_into_tensor on every build
        parts = [torch.zeros(maxB, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(parts, send[:maxB].clone())
        v[n_own:n_own + self.world * maxB] = torch.cat(parts)


def main():
    name, n_dist, out, halo = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ml, ex = load_hierarchy(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    be = EmuBackend(rank, world)
    ds = D.DistributedSolver(ml, be, n_dist=n_dist, halo=halo)
    ds.load(ex["b"])
    ncyc = 3
    norms = be.vector(ncyc + 1)
    ds.cycles(ncyc, norms=norms)
    x = ds.gather_x()
    if rank == 0:
        np.savez(out, x=x, res=np.sqrt(norms[:ncyc + 1].numpy()), launches=np.array([be.kernel_launches]))
    dist.barrier()
    be.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
