"""Parity at BASELINE.json's FULL sizes through size-independent properties.

The CPU oracle cannot run a 16384 x 16384 lattice inside a test, but the
lattice Boltzmann step is strictly local (a node at step N depends on data at
Chebyshev distance <= N), so a full-size field is completely determined by a
small oracle run:

* configs[3], lid-driven cavity 16384^2, rho = 1, u = 0: after N steps every
  node closer than S/2 to a corner equals the same node of an S x S cavity,
  every node near one wall only equals the mid-wall node of the small cavity
  and everything else equals its centre node (``locality``);
* configs[4], x-periodic channel 8192 x 16384 (one GPU's slab), MRT + Guo:
  an initial state with period P along x stays P-periodic, so the field is
  the P x 16384 oracle run tiled along x (``translation invariance``); the
  total mass is conserved (``checksum``).

BGK + strict build: bit-exact.  Production build / MRT: 1e-12 relative
(max-norm per field), the tolerance BASELINE.json states.
"""
import numpy as np
import pytest

import cases
from pylabolt_b200 import capi
from test_gpu_parity import RTOL, make_solver, oracle_for

pytestmark = pytest.mark.gpu

FULL = 16384


def _need_memory(gb):
    import subprocess
    try:
        out = subprocess.check_output(
            ["nvidia-smi", "--query-gpu=memory.total",
             "--format=csv,noheader,nounits", "-i", "0"], text=True)
        if float(out.strip().splitlines()[0]) < gb * 1024:
            pytest.skip(f"needs {gb} GB of device memory")
    except (OSError, subprocess.CalledProcessError, ValueError):
        pass


def _fold(n_full, n_small):
    """Index map full -> small: the first and last n_small/2 nodes map to
    themselves (counted from their own wall), everything between to the
    centre node of the small lattice."""
    half = n_small // 2
    i = np.arange(n_full)
    return np.where(i < half, i,
                    np.where(i >= n_full - half, i - (n_full - n_small), half))


def _inner(flat, nx, ny, ncomp=1):
    a = flat.reshape(nx + 2, ny + 2, ncomp)[1:-1, 1:-1]
    return a[..., 0] if ncomp == 1 else a


@pytest.mark.parametrize("strict", [True, False])
def test_cavity_16384_is_determined_by_a_small_cavity(strict):
    """configs[3] at full size against the oracle through locality."""
    _need_memory(60)
    steps, small = 24, 128          # small/2 = 64 > steps + 2
    s = make_solver(cases.cavity(FULL, FULL, end_time=steps), strict=strict)
    ref = make_solver(cases.cavity(small, small, end_time=steps), strict=strict)
    try:
        orc = oracle_for(ref, n_threads=8)
        orc.step(steps)
        info = s.plb.info()
        assert info["n_bulk"] + info["n_link"] == FULL * FULL
        s.advance(steps, store_moments_last=True)
        mx = _fold(FULL, small)
        for field, ncomp, want_flat in ((capi.DENSITY, 1, orc.density),
                                        (capi.VELOCITY, 2, orc.velocity)):
            got = _inner(s.plb.download(field), FULL, FULL, ncomp)
            want = _inner(want_flat, small, small, ncomp)
            scale = np.abs(want).max()
            worst = 0.0
            for x0 in range(0, FULL, 1024):      # bounded temporaries
                rows = mx[x0:x0 + 1024]
                expect = want[rows][:, mx]
                block = got[x0:x0 + 1024]
                if strict:
                    assert np.array_equal(block, expect), (field, x0)
                else:
                    worst = max(worst, float(np.abs(block - expect).max()))
            assert worst <= RTOL * scale
            del got
    finally:
        s.close()
        ref.close()


def _periodic_noise(period, amplitude):
    """Integer hash noise with period `period` along x (no transcendental
    functions: the value of a node must not depend on where it sits in a
    vectorised numpy call)."""
    def func(i, j):
        i = np.asarray(i, dtype=np.int64) % period
        j = np.asarray(j, dtype=np.int64)
        h = (i * 73856093) ^ (j * 19349663)
        h = (h ^ (h >> 13)) * 1274126177
        a = ((h >> 8) & 0xFFFF).astype(np.float64) / 65536.0 - 0.5
        b = ((h >> 24) & 0xFFFF).astype(np.float64) / 65536.0 - 0.5
        return amplitude * a, amplitude * b
    func.vectorized = True
    return func


def _channel(nx, ny, model, steps, period):
    sim = cases.poiseuille(nx, ny, end_time=steps, forcing="guo_second_order",
                           g=1.0e-6, kin_visc=0.1, model=model)
    fluid = sim.initial_fields_dict["default"]["fluid"]
    fluid["velocity"] = {"type": "func", "func": _periodic_noise(period, 0.01)}
    fluid["density"] = {"type": "fixed", "value": 1.0}
    return sim


@pytest.mark.parametrize("model,strict", [("MRT", False), ("BGK", True)])
def test_channel_8192x16384_is_the_tiled_periodic_strip(model, strict):
    """configs[4] (one GPU's slab, with the x-periodic seam) at full size:
    a P-periodic state equals the P x ny oracle run tiled along x."""
    _need_memory(30)
    nx, ny, period, steps = 8192, FULL, 32, 20
    s = make_solver(_channel(nx, ny, model, steps, period), strict=strict)
    ref = make_solver(_channel(period, ny, model, steps, period), strict=strict)
    try:
        orc = oracle_for(ref, n_threads=8)
        mass0 = float(_inner(orc.density, period, ny).sum()) * (nx // period)
        orc.step(steps)
        s.advance(steps, store_moments_last=True)
        mass = 0.0
        for field, ncomp, want_flat in ((capi.DENSITY, 1, orc.density),
                                        (capi.VELOCITY, 2, orc.velocity)):
            got = _inner(s.plb.download(field), nx, ny, ncomp)
            want = _inner(want_flat, period, ny, ncomp)
            if field == capi.DENSITY:
                mass = float(got.sum())
            scale = np.abs(want).max()
            for x0 in range(0, nx, 1024):        # bounded temporaries
                block = got[x0:x0 + 1024]
                tiles = block.reshape((-1, period) + block.shape[1:])
                if strict:
                    assert np.array_equal(
                        tiles, np.broadcast_to(want, tiles.shape)), (field, x0)
                else:
                    assert float(np.abs(tiles - want).max()) <= RTOL * scale
            del got
        # closed channel (periodic + bounce back): mass is conserved
        assert abs(mass - mass0) <= 1e-13 * mass0
    finally:
        s.close()
        ref.close()
