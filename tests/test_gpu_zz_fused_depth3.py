"""PLB_FUSE_DEPTH=3 on the GPU: three steps per pass (opt-in until it has been
timed).  Same kernel template as the two-step path, one
hand-over deeper.  Kept in its own module, collected after the other GPU
modules, because these cases had not run on a B200 when round 1 ended.
"""
import numpy as np
import pytest

from test_emu_parity import WIDE_CASES
from test_gpu_fused import MID_CASES, _fields

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(WIDE_CASES))
def test_three_steps_per_pass_equals_single_steps(name, monkeypatch):
    """PLB_FUSE_DEPTH=3 (opt-in): 14 plain steps = 4 triples + 1 pair."""
    factory = WIDE_CASES[name]
    want, _ = _fields(factory, 15, "0", True, monkeypatch)
    got, info = _fields(factory, 15, "2", True, monkeypatch, depth=3)
    if info["n_deep3"] > 0:
        assert info["triples"] == 4 and info["pairs"] == 1, info
    for key in ("density", "velocity", "pop_fluid_new"):
        assert np.array_equal(got[key], want[key]), key


@pytest.mark.parametrize("name", sorted(MID_CASES))
def test_three_steps_per_pass_mid_size(name, monkeypatch):
    factory = MID_CASES[name]
    want, _ = _fields(factory, 21, "0", True, monkeypatch)
    got, info = _fields(factory, 21, "1", True, monkeypatch, depth=3)
    assert info["active"] == 3 and info["triples"] == 6 and info["pairs"] == 1
    for key in ("density", "velocity", "pop_fluid_new"):
        assert np.array_equal(got[key], want[key]), key
