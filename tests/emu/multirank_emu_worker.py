"""TEST INFRASTRUCTURE: several ranks of libplb in ONE process, on the CPU.

Each rank is a host thread that drives its own solver through the emulated
library (tests/emu/README.md), exactly like one process per GPU would: the
emulated streams queue their operations and a rank that waits for its
neighbour (mailbox flag of the peer-to-peer faces, or a receive) simply stays
blocked until the neighbour's thread has issued the matching work.  CUDA IPC
handles are plain pointers and the NCCL calls libplb makes are served by an
in-process stand-in (emu_runtime.cpp, reached through a libnccl.so.2 shim on
LD_LIBRARY_PATH).  The decomposed result is compared with the CPU oracle on the
undecomposed domain by tests/multirank_worker.run_one -- the function the
multi-GPU test runs on the box.

    python tests/emu/multirank_emu_worker.py <world> <out.json> <spec> [...]
"""
import json
import os
import sys
import tempfile
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
REPO = os.path.dirname(TESTS)
for p in (REPO, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)


from pylabolt_b200 import capi  # noqa: E402
capi._accept_emulated_build = True      # this worker IS the emulation harness


class ThreadWorld:
    def __init__(self, size):
        self.size = size
        self.barrier = threading.Barrier(size)
        self.slots = [None] * size
        self.box = None


class ThreadComm:
    """The solver's communicator surface (pylabolt_b200/comm.py) over threads."""

    def __init__(self, world, rank):
        self.world, self.rank = world, rank

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.world.size

    def Barrier(self):
        self.world.barrier.wait()

    def Abort(self, code=1):
        pass

    def Allreduce(self, local, out, op="sum"):
        w = self.world
        w.slots[self.rank] = np.array(local, copy=True)
        w.barrier.wait()
        stack = np.stack(w.slots)
        red = {"sum": stack.sum(axis=0), "max": stack.max(axis=0),
               "min": stack.min(axis=0)}[op]
        w.barrier.wait()
        out[...] = red

    def bcast_bytes(self, data, root=0):
        w = self.world
        if self.rank == root:
            w.box = bytes(data)
        w.barrier.wait()
        got = w.box
        w.barrier.wait()
        return got


def main():
    world_size, out_path, specs = int(sys.argv[1]), sys.argv[2], sys.argv[3:]
    import multirank_worker as mw        # run_one, CASES (no torch import)
    world = ThreadWorld(world_size)
    results = {}
    failures = []

    def rank_main(rank, tmp):
        comm = ThreadComm(world, rank)
        try:
            for spec in specs:
                name, steps, mode, face, *rest = spec.split(":")
                fuse, depth = mw.fuse_options(rest)
                status = mw.run_one(comm, name, int(steps), mode == "strict", face,
                                    tmp, fuse=fuse, depth=depth)
                if rank == 0:
                    results[spec] = status
        except BaseException as e:       # noqa: BLE001 -- the other ranks would hang
            failures.append(f"rank {rank}: {type(e).__name__}: {e}")
            world.barrier.abort()

    with tempfile.TemporaryDirectory() as tmp:
        threads = [threading.Thread(target=rank_main, args=(r, tmp))
                   for r in range(world_size)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    for spec in specs:
        results.setdefault(spec, 2)
    with open(out_path, "w") as f:
        json.dump({"results": results, "failures": failures}, f)
    sys.exit(max(results.values()) if not failures else 3)


if __name__ == "__main__":
    main()
