"""bench.py's own parity check (ParityCheck) checked on the CPU: the small
oracle run it folds out / tiles must reproduce an oracle run of the WHOLE
lattice bit for bit (locality + periodic initial pattern), per slab."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def _whole(workload, nx, ny, steps):
    sim = bench.WORKLOADS[workload][0](nx, ny, 1, 1)
    orc, _ = bench.make_oracle(sim, 4)
    orc.step(steps)
    rho = orc.density.reshape(nx + 2, ny + 2)[1:-1, 1:-1]
    u = orc.velocity.reshape(nx + 2, ny + 2, 2)[1:-1, 1:-1]
    return rho, u


@pytest.mark.parametrize("workload,nx,ny,steps", [
    ("cavity", 384, 320, 13),       # small cavity 2 * 32 + 32 = 96 per direction
    ("cavity", 64, 64, 9),          # too small to fold: oracle runs the whole case
    ("channel", 256, 96, 17)])
def test_small_oracle_run_determines_the_whole_lattice(workload, nx, ny, steps):
    rho, u = _whole(workload, nx, ny, steps)
    for n_slabs in (1, 3):
        chunk = -(-nx // n_slabs)
        for r in range(n_slabs):
            x0 = r * chunk
            nxr = min(chunk, nx - x0)
            chk = bench.ParityCheck(workload, nx, ny, x0, nxr, 4)
            want = chk.expected(steps)
            if workload == "cavity" and nx > 300:
                assert chk.small == (96, 96)
            err = chk.compare(want, np.ascontiguousarray(rho[x0:x0 + nxr]),
                              np.ascontiguousarray(u[x0:x0 + nxr]))
            assert err == 0.0
            # ... and it notices a single wrong node
            bad = np.ascontiguousarray(rho[x0:x0 + nxr]).copy()
            bad[nxr // 2, ny // 3] += 1e-9
            assert chk.compare(want, bad,
                               np.ascontiguousarray(u[x0:x0 + nxr])) > 1e-10


def test_expected_is_incremental():
    chk = bench.ParityCheck("channel", 64, 40, 0, 64, 2)
    a = chk.expected(4, steps_max=10)
    b = chk.expected(10)
    chk2 = bench.ParityCheck("channel", 64, 40, 0, 64, 2)
    c = chk2.expected(10)
    assert np.array_equal(b[0], c[0]) and np.array_equal(b[1], c[1])
    assert not np.array_equal(a[0], b[0])
