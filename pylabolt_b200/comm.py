"""Communicators with the small mpi4py-like surface the solver uses
(Get_rank / Get_size / Barrier / Abort / Allreduce / bcast_bytes).

The reference talks to mpi4py (pylabolt/parallel/MPI_operator.py); the b200
back end runs one process per GPU under torchrun and uses torch.distributed
purely as plumbing: rendezvous, the broadcast of the NCCL unique id that
libplb's own communicator is created from, and the few-double reductions of
residues / forces.  Lattice data never goes through this class -- the slab
faces are exchanged inside libplb.
"""
import os

import numpy as np


class SingleComm:
    """One rank, no distributed runtime."""

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Barrier(self):
        pass

    def Abort(self, code=1):
        """The caller re-raises the configuration error."""

    def Allreduce(self, local, out, op="sum"):
        out[...] = local

    def bcast_bytes(self, data, root=0):
        return data


class TorchComm:
    """torch.distributed process group (nccl on GPUs, gloo on CPU)."""

    def __init__(self, backend=None, init=True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        if init and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend == "nccl":
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group(backend=backend)
        self.backend = dist.get_backend()
        self.device = (torch.device("cuda", torch.cuda.current_device())
                       if self.backend == "nccl" else torch.device("cpu"))

    def Get_rank(self):
        return self.dist.get_rank()

    def Get_size(self):
        return self.dist.get_world_size()

    def Barrier(self):
        self.dist.barrier()

    def Abort(self, code=1):
        """Configuration errors are raised on every rank (all ranks parse the
        same case file), so the job ends without a collective teardown."""

    def Allreduce(self, local, out, op="sum"):
        t = self.torch.as_tensor(np.ascontiguousarray(local)).to(self.device)
        red = {"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX,
               "min": self.dist.ReduceOp.MIN}[op]
        self.dist.all_reduce(t, op=red)
        out[...] = t.cpu().numpy()

    def bcast_bytes(self, data, root=0):
        n = len(data)
        if self.Get_rank() == root:
            t = self.torch.tensor(list(data), dtype=self.torch.uint8)
        else:
            t = self.torch.zeros(n, dtype=self.torch.uint8)
        t = t.to(self.device)
        self.dist.broadcast(t, src=root)
        return bytes(t.cpu().numpy().tobytes())


def world_comm():
    """TorchComm under torchrun (WORLD_SIZE > 1), SingleComm otherwise."""
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return TorchComm()
    return SingleComm()
