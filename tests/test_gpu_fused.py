"""The two-steps-per-pass path (k_bulk_fused / step_fused) on the GPU.

Per node and step the fused path runs the same collide<>() as the single-step
kernels, so in the -fmad=false build it must reproduce them BIT FOR BIT, and
the reference's golden vectors with them; the production build stays within
the 1e-12 of BASELINE.json.  Sizes here are chosen so that whole 62-node warp
strips are deep (128-bit shuffle-assembled stores), several strips and row
chunks exist, and the list passes, the periodic seam and the zero_gradient
pass interleave with the fused kernel on two streams.
"""
import os

import numpy as np
import pytest

import cases
from pylabolt_b200 import capi
from test_emu_parity import WIDE_CASES, _mrt, oracle_for, rel_err
from test_gpu_parity import make_solver

pytestmark = pytest.mark.gpu

RTOL = 1e-12


@pytest.fixture(autouse=True)
def two_steps_per_pass(monkeypatch):
    """This module counts PAIRS: it pins two steps per pass.  The shipped
    default (three, since round 2) is covered by test_gpu_zz_fused_depth3.py,
    the full-size tests and every other GPU module."""
    monkeypatch.setenv("PLB_FUSE_DEPTH", "2")


def _fields(sim_factory, n_steps, fuse, strict, monkeypatch, one_by_one=False,
            depth=2):
    monkeypatch.setenv("PLB_FUSE", fuse)
    monkeypatch.setenv("PLB_FUSE_DEPTH", str(depth))
    s = make_solver(sim_factory(), strict=strict)
    try:
        if one_by_one:
            for _ in range(n_steps - 1):
                s.execute_single_time_step()
            s.single_time_step(store_moments=True)
        else:
            s.advance(n_steps, store_moments_last=True)
        return s.fields_to_host(), s.plb.fused_info()
    finally:
        s.close()


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CASES))
def test_fused_strict_build_is_bit_exact_with_reference(golden_dir, name,
                                                        monkeypatch):
    monkeypatch.setenv("PLB_FUSE", "2")
    factory, kwargs, record = cases.GOLDEN_CASES[name]
    data = np.load(os.path.join(golden_dir, name + ".npz"))
    s = make_solver(factory(**kwargs), strict=True)
    try:
        done = 0
        for step in record:
            s.advance(step - done, store_moments_last=True)
            done = step
            got = s.fields_to_host()
            assert np.array_equal(got["density"], data[f"density_{step}"]), step
            assert np.array_equal(got["velocity"], data[f"velocity_{step}"]), step
            assert np.array_equal(got["pop_fluid_new"], data[f"pop_{step}"]), step
        assert s.plb.fused_info()["pairs"] > 0
    finally:
        s.close()


@pytest.mark.parametrize("name", sorted(WIDE_CASES))
def test_fused_equals_single_steps(name, monkeypatch):
    factory = WIDE_CASES[name]
    want, _ = _fields(factory, 13, "0", True, monkeypatch)
    got, info = _fields(factory, 13, "2", True, monkeypatch)
    assert info["pairs"] == (6 if info["n_deep"] else 0)
    for key in ("density", "velocity", "pop_fluid_new"):
        assert np.array_equal(got[key], want[key]), key
    # production build (FMA contraction may differ between the two kernels)
    want, _ = _fields(factory, 13, "0", False, monkeypatch)
    got, _ = _fields(factory, 13, "2", False, monkeypatch, one_by_one=True)
    for key in ("density", "velocity", "pop_fluid_new"):
        assert rel_err(got[key], want[key]) <= RTOL, key


MID_CASES = {
    # many strips / chunks / CTAs: the grid of the fused kernel is > 1 wave
    "channel_mrt_guo2_900x1300": lambda: _mrt(cases.poiseuille(900, 1300)),
    "cavity_bgk_1111x1300": lambda: cases.cavity(1111, 1300),
    "cylinder_bgk_800x600": lambda: cases.cylinder(800, 600, radius=80),
}


@pytest.mark.parametrize("name", sorted(MID_CASES))
def test_fused_equals_single_steps_mid_size(name, monkeypatch):
    factory = MID_CASES[name]
    want, _ = _fields(factory, 21, "0", True, monkeypatch)
    got, info = _fields(factory, 21, "1", True, monkeypatch)   # default mode
    assert info["active"] == 2 and info["pairs"] == 10
    for key in ("density", "velocity", "pop_fluid_new"):
        assert np.array_equal(got[key], want[key]), key


@pytest.mark.parametrize("name", ["poiseuille_70x140_guo2", "cylinder_120x140",
                                  "mrt_poiseuille_70x140_guo2",
                                  "zero_gradient_100x127"])
@pytest.mark.parametrize("strict", [True, False])
def test_fused_against_oracle(name, strict, monkeypatch):
    monkeypatch.setenv("PLB_FUSE", "2")
    s = make_solver(WIDE_CASES[name](), strict=strict)
    try:
        orc = oracle_for(s)
        for n in (2, 7, 40):
            s.advance(n, store_moments_last=True)
            orc.step(n)
            got = s.fields_to_host()
            assert rel_err(got["density"], orc.density) <= RTOL
            assert rel_err(got["velocity"], orc.velocity) <= RTOL
            assert rel_err(got["pop_fluid_new"], orc.pop_new) <= RTOL
            bgk = s.collision_operator.collision_fluid == "BGK"
            if strict and bgk and "zero_gradient" not in name:
                assert np.array_equal(got["pop_fluid_new"], orc.pop_new)
        assert s.plb.fused_info()["pairs"] == 0 + 3 + 19
    finally:
        s.close()


def test_mass_is_conserved_with_pairing(monkeypatch):
    monkeypatch.setenv("PLB_FUSE", "2")
    sim = cases.periodic_box(256, 192, forcing=None)
    sim.obstacle_dict = {"options": {}}
    s = make_solver(sim, strict=False)
    try:
        s.advance(1, store_moments_last=True)
        m0 = s.plb.download(capi.DENSITY_INNER).sum()
        s.advance(401, store_moments_last=True)
        m1 = s.plb.download(capi.DENSITY_INNER).sum()
        assert s.plb.fused_info()["pairs"] == 200
        assert abs(m1 - m0) <= 1e-11 * abs(m0)
    finally:
        s.close()
