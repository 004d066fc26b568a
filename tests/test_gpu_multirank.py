"""x-slab decomposition across GPUs (one process per GPU, NCCL face exchange
inside libplb) against the CPU oracle on the undecomposed domain.  Needs at
least two GPUs: `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multirank`.
"""
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


def run_case(name, world, steps, mode, port):
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port),
               os.path.join(HERE, "multirank_worker.py"), name, str(steps),
               mode, tmp]
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        print(proc.stdout[-3000:])
        print(proc.stderr[-3000:])
        return proc.returncode


@pytest.mark.skipif(N_GPUS < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", ["strict", "production"])
@pytest.mark.parametrize("name", ["cavity", "poiseuille", "cylinder_cut",
                                  "periodic_box", "mrt_box", "spin"])
def test_two_slabs_match_oracle(name, mode):
    assert run_case(name, 2, 25, mode, 29641) == 0


@pytest.mark.skipif(N_GPUS < 4, reason="needs >= 4 GPUs")
@pytest.mark.parametrize("name", ["poiseuille", "cylinder_cut", "periodic_box"])
def test_four_slabs_match_oracle(name):
    assert run_case(name, 4, 25, "strict", 29642) == 0
