"""world_size-2 (and 3) runs on CPU over gloo: the host side of the multi-GPU
path -- communicator plumbing, slab ownership, ghost-column flags -- driven by
real processes.  See tests/gloo_worker.py."""
import os
import subprocess
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def run(name, world, steps, port):
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port),
               os.path.join(HERE, "gloo_worker.py"), name, str(steps), tmp]
        env = dict(os.environ, OMP_NUM_THREADS="2", CUDA_VISIBLE_DEVICES="")
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                              env=env)
        sys.stdout.write(proc.stdout[-2000:])
        sys.stderr.write(proc.stderr[-2000:])
        return proc.returncode


@pytest.mark.parametrize("name,world", [("cavity", 2), ("poiseuille", 2),
                                        ("periodic_box", 2), ("spin_cut", 2),
                                        ("periodic_box", 3)])
def test_slabs_over_gloo_equal_undecomposed_domain(name, world):
    assert run(name, world, 20, 29700 + world) == 0
