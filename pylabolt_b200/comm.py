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

    def all_agree(self, flag):
        return bool(flag)

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
        """MPI_Abort semantics (the reference calls comm.Abort() on every
        fatal error): the calling rank leaves at once with a non-zero status,
        without waiting in any collective; torchrun then terminates the other
        ranks, and a peer that is already spinning on this rank's slab face
        gives up after PLB_P2P_TIMEOUT_S.  A rank-local failure (a missing
        device, overlapping obstacles seen by one slab only) therefore ends
        the job instead of hanging it."""
        import sys
        sys.stdout.flush()
        sys.stderr.flush()
        if self.Get_size() > 1:
            os._exit(code if code else 1)

    def all_agree(self, flag):
        """True on every rank iff `flag` is true on every rank."""
        mine = np.array([1.0 if flag else 0.0])
        out = np.zeros_like(mine)
        self.Allreduce(mine, out, op="min")
        return bool(out[0] > 0.5)

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
