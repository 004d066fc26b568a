"""x-slab decomposition across GPUs (one process per GPU) against the CPU
oracle on the undecomposed domain, with both face transports of libplb:
peer-to-peer stores over NVLink (the default) and NCCL send/recv
(PLB_FACE=nccl).  Needs at least two GPUs:
`gpurun --gpus 2 -- python -m pytest tests -m gpu -k multirank`.

All cases of one world size run inside ONE torchrun launch (process start-up
and NCCL initialisation dominate a case); each test then looks up its own
result.
"""
import json
import os
import subprocess
import sys
import tempfile

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True,
                             text=True, timeout=30).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


N_GPUS = _gpu_count()

NAMES = ["cavity", "poiseuille", "cylinder_cut", "periodic_box", "mrt_box",
         "spin"]
# every case on both transports; the strict (bit-exact) build on the default
# transport, plus two long runs that give a missed hand-shake time to show
TWO_SLABS = ([f"{n}:25:strict:p2p" for n in NAMES] +
             [f"{n}:25:production:nccl" for n in NAMES] +
             ["periodic_box:25:production:p2p", "cylinder_cut:25:strict:nccl",
              "periodic_box:400:strict:p2p", "poiseuille:400:strict:nccl",
              "thin:60:strict:p2p", "thin:60:strict:nccl"] +
             # two steps per pass (PLB_FUSE=2): both transports, a body on the
             # cut, the periodic seam through the ring, long runs
             [f"{n}:25:strict:p2p:fuse2" for n in NAMES] +
             ["cylinder_cut:26:strict:nccl:fuse2",
              "periodic_box:400:strict:p2p:fuse2",
              "poiseuille:401:strict:nccl:fuse2",
              "wide_channel:41:production:p2p:fuse2",
              "wide_cylinder:40:strict:p2p:fuse2",
              "wide_channel:40:strict:nccl:fuse2"] +
             # three steps per pass (PLB_FUSE_DEPTH=3, opt-in)
             ["periodic_box:25:strict:p2p:fuse3", "cylinder_cut:26:strict:nccl:fuse3",
              "wide_channel:41:strict:p2p:fuse3", "wide_cylinder:40:strict:nccl:fuse3",
              "poiseuille:400:strict:p2p:fuse3"] +
             # four steps per pass
             ["wide_channel:42:strict:p2p:fuse4", "wide_cylinder:41:strict:nccl:fuse4",
              "poiseuille:401:strict:p2p:fuse4"])
FOUR_SLABS = ([f"{n}:25:strict:p2p" for n in
               ("poiseuille", "cylinder_cut", "periodic_box")] +
              ["periodic_box:25:strict:nccl", "mrt_box:300:production:p2p",
               "thin:60:strict:p2p", "thin:60:strict:nccl",
               "uneven:40:strict:p2p", "uneven:41:strict:p2p:fuse2",
               "periodic_box:25:strict:p2p:fuse2",
               "wide_channel:60:strict:nccl:fuse2",
               "wide_cylinder:60:strict:p2p:fuse2",
               "uneven:41:strict:p2p:fuse3", "wide_cylinder:61:strict:p2p:fuse3",
               "wide_channel:61:strict:p2p:fuse4"])


def _launch(world, specs, port):
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port),
               os.path.join(HERE, "multirank_worker.py"), tmp] + specs
        env = dict(os.environ, PLB_P2P_TIMEOUT_S="20")
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                              env=env)
        log = proc.stdout[-6000:] + "\n" + proc.stderr[-3000:]
        try:
            with open(os.path.join(tmp, "results.json")) as f:
                return json.load(f), log
        except OSError:
            return {}, log


@pytest.fixture(scope="module")
def two_slabs():
    return _launch(2, TWO_SLABS, 29641)


@pytest.fixture(scope="module")
def four_slabs():
    return _launch(4, FOUR_SLABS, 29642)


@pytest.mark.skipif(N_GPUS < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("spec", TWO_SLABS)
def test_two_slabs_match_oracle(two_slabs, spec):
    results, log = two_slabs
    assert results.get(spec) == 0, log


@pytest.mark.skipif(N_GPUS < 4, reason="needs >= 4 GPUs")
@pytest.mark.parametrize("spec", FOUR_SLABS)
def test_four_slabs_match_oracle(four_slabs, spec):
    results, log = four_slabs
    assert results.get(spec) == 0, log
