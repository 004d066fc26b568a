"""Three steps per pass on the GPU (the shipped default for BGK; the
two-stress-moment MRT kernel takes four, test_gpu_zz_fused_depth4.py;
DESIGN.md section 3a).  Same kernel template as the two-step path, one
hand-over deeper.
"""
import numpy as np
import pytest

from test_emu_parity import WIDE_CASES
from test_gpu_fused import MID_CASES, _fields

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(WIDE_CASES))
def test_three_steps_per_pass_equals_single_steps(name, monkeypatch):
    """14 plain steps = 4 triples + 1 pair."""
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


def test_default_steps_per_pass(monkeypatch):
    """No environment at all: the library groups plain steps four at a time
    for the two-stress-moment MRT kernel and three at a time for BGK, and the
    result is that of single steps, bit for bit (strict build)."""
    monkeypatch.delenv("PLB_FUSE_DEPTH", raising=False)
    monkeypatch.delenv("PLB_FUSE", raising=False)
    from test_gpu_parity import make_solver
    for name, depth, groups in (("channel_mrt_guo2_900x1300", 4, {"quads": 5}),
                                ("cavity_bgk_1111x1300", 3, {"triples": 6, "pairs": 1})):
        factory = MID_CASES[name]
        s = make_solver(factory(), strict=True)
        try:
            assert s.plb.fused_info()["active"] == depth
            s.advance(20)
            s.advance(1, store_moments_last=True)
            info = s.plb.fused_info()
            for key in ("pairs", "triples", "quads"):
                assert info[key] == groups.get(key, 0), (name, info)
            got = s.fields_to_host()
            assert "ring=tma-tensor carry=shared" in s.plb.build_info()
        finally:
            s.close()
        want, _ = _fields(factory, 21, "0", True, monkeypatch)
        monkeypatch.delenv("PLB_FUSE_DEPTH", raising=False)
        monkeypatch.delenv("PLB_FUSE", raising=False)
        for key in ("density", "velocity", "pop_fluid_new"):
            assert np.array_equal(got[key], want[key]), (name, key)
