"""Four steps per pass on the GPU (PLB_FUSE_DEPTH=4):
the same kernel template one more hand-over deeper (58 of 64 nodes per warp
strip, chunks overlap by six rows, deep flags up to 3, four list passes, three
scratch lattices).  Strict build, bit for bit against single steps.
"""
import numpy as np
import pytest

from test_emu_parity import WIDE_CASES, random_bodies_case
from test_gpu_fused import MID_CASES, _fields

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(WIDE_CASES))
def test_four_steps_per_pass_equals_single_steps(name, monkeypatch):
    """14 plain steps = 3 groups of four + 1 pair."""
    factory = WIDE_CASES[name]
    want, _ = _fields(factory, 15, "0", True, monkeypatch)
    got, info = _fields(factory, 15, "2", True, monkeypatch, depth=4)
    if info["n_deep4"] > 0:
        assert info["quads"] == 3 and info["pairs"] == 1, info
    for key in ("density", "velocity", "pop_fluid_new"):
        assert np.array_equal(got[key], want[key]), key


@pytest.mark.parametrize("name", sorted(MID_CASES))
def test_four_steps_per_pass_mid_size(name, monkeypatch):
    """20 plain steps = 5 groups of four; 900 x 1300 and the like: many chunks
    and strips, fully deep warps (the 128-bit store path)."""
    factory = MID_CASES[name]
    want, _ = _fields(factory, 21, "0", True, monkeypatch)
    got, info = _fields(factory, 21, "1", True, monkeypatch, depth=4)
    assert info["active"] == 4 and info["quads"] == 5 and info["pairs"] == 0, info
    for key in ("density", "velocity", "pop_fluid_new"):
        assert np.array_equal(got[key], want[key]), key


def test_remainder_of_three_is_one_three_step_pass(monkeypatch):
    """23 plain steps at depth 4 = 5 groups of four + one of three."""
    factory = MID_CASES[sorted(MID_CASES)[0]]
    want, _ = _fields(factory, 24, "0", True, monkeypatch)
    got, info = _fields(factory, 24, "1", True, monkeypatch, depth=4)
    assert info["quads"] == 5 and info["triples"] == 1 and info["pairs"] == 0, info
    for key in ("density", "velocity", "pop_fluid_new"):
        assert np.array_equal(got[key], want[key]), key


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_bodies_every_depth_equals_single_steps(seed, monkeypatch):
    """Randomly placed circles / ellipses in a 400 x 520 channel: deep flags,
    list passes and the fused kernel of every depth add up to single steps."""
    factory = lambda: random_bodies_case(seed, 400, 520)
    want, _ = _fields(factory, 14, "0", True, monkeypatch)
    for depth in (2, 3, 4):
        got, info = _fields(factory, 14, "2", True, monkeypatch, depth=depth)
        assert info["pairs"] + info["triples"] + info["quads"] > 0, info
        for key in ("density", "velocity", "pop_fluid_new"):
            assert np.array_equal(got[key], want[key]), (seed, depth, key)
