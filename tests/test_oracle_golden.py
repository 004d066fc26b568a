"""Pins the CPU oracle (oracle/plb_oracle.c) to the reference: every fixture in
tests/golden/ was produced by the reference's own numba kernels
(tests/golden/make_golden.py); the oracle must reproduce rho, u and pop_new
BIT FOR BIT on all of them (all golden cases are BGK paths)."""
import os

import numpy as np
import pytest

import cases
from oracle.oracle import Oracle, elements_from_golden


def load_golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def oracle_from_golden(data, simulation, **kw):
    fluid = simulation.collision_dict["fluid"]
    return Oracle(
        data["shape"], data["solid"], data["ghost_node"], data["density_0"],
        data["velocity_0"], elements_from_golden(data), float(data["omega"]),
        gravity=data["gravity"], forcing=fluid["forcing_model"],
        collision=fluid["model"], x_periodic=bool(data["x_periodic"]),
        y_periodic=bool(data["y_periodic"]), **kw)


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CASES))
def test_oracle_matches_reference_bit_for_bit(golden_dir, name):
    factory, kwargs, record = cases.GOLDEN_CASES[name]
    data = load_golden(golden_dir, name)
    orc = oracle_from_golden(data, factory(**kwargs), n_threads=2)
    orc.initialize_pop()
    assert np.array_equal(orc.pop_new, data["pop_0"])
    done = 0
    for step in record:
        orc.step(step - done)
        done = step
        assert np.array_equal(orc.density, data[f"density_{step}"]), step
        assert np.array_equal(orc.velocity, data[f"velocity_{step}"]), step
        assert np.array_equal(orc.pop_new, data[f"pop_{step}"]), step


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CASES))
def test_oracle_residues_match_the_reference_operator(golden_dir, name):
    """oracle_residue_sums against the reference's own ResidueOperator
    (utils/residues.py:171-222 + cpu/compute_residues_kernels.py:6-73) called
    at the recorded steps: the six sums it reduces and the three residues it
    logs.  The reference sums with a numba prange reduction whose order is
    not fixed, hence a rounding-level tolerance instead of bit equality."""
    factory, kwargs, record = cases.GOLDEN_CASES[name]
    data = load_golden(golden_dir, name)
    orc = oracle_from_golden(data, factory(**kwargs), n_threads=2)
    orc.initialize_pop()
    done = 0
    for step in record:
        orc.step(step - done)
        done = step
        res = orc.residues()
        sums = orc.last_residue_sums
        assert np.allclose(sums, data[f"residue_sums_{step}"], rtol=1e-12,
                           atol=1e-300), step
        assert np.allclose(res, data[f"residues_{step}"], rtol=1e-12,
                           atol=1e-300), step


def test_lattice_constants_match_reference(golden_dir):
    """base/lattice.py:41-60 values, as probed from the reference."""
    from oracle.oracle import lattice_constants
    data = load_golden(golden_dir, "cavity")
    c = lattice_constants()
    ref = data["lattice_consts"]
    assert c["cs"] == ref[0] and c["cs_2"] == ref[1]
    assert c["inv_cs_2"] == ref[2] == 2.999999999999999
    assert c["inv_cs_4"] == ref[3]
    assert c["float_min"] == ref[4] == 2.220446049250313e-16
    assert np.array_equal(c["weights"], data["weights"])
    assert float(data["omega"]) == 1.25


def test_thread_count_does_not_change_results(golden_dir):
    """Hot kernels have no cross-node reductions."""
    factory, kwargs, _ = cases.GOLDEN_CASES["cylinder_spin"]
    data = load_golden(golden_dir, "cylinder_spin")
    a = oracle_from_golden(data, factory(**kwargs), n_threads=1)
    a.initialize_pop()
    a.step(25)
    b = oracle_from_golden(data, factory(**kwargs), n_threads=4)
    b.initialize_pop()
    b.step(25)
    assert np.array_equal(a.pop_new, b.pop_new)


def test_mrt_with_uniform_rates_reduces_to_bgk(golden_dir):
    """Our MRT definition with S = omega * 1 is
    algebraically the BGK + Guo update; tolerance is rounding only."""
    factory, kwargs, _ = cases.GOLDEN_CASES["periodic_box"]
    data = load_golden(golden_dir, "periodic_box")
    sim = factory(**kwargs)
    omega = float(data["omega"])
    bgk = oracle_from_golden(data, sim)
    bgk.initialize_pop()
    bgk.step(20)
    sim.collision_dict["fluid"]["model"] = "MRT"
    mrt = oracle_from_golden(data, sim, mrt_rates=[omega] * 9)
    mrt.initialize_pop()
    mrt.step(20)
    scale = np.abs(bgk.pop_new).max()
    assert np.abs(mrt.pop_new - bgk.pop_new).max() <= 1e-13 * scale
