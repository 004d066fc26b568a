"""Host-side setup of pylabolt_b200 against the reference: node flags, solid
ids, boundary link lists, surface normals and the initial rho / u fields must
be BIT-EXACT with what the reference's own containers produced
(tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import cases
from pylabolt_b200.comm import SingleComm
from pylabolt_b200.operators import CollisionOperator, FluidLB, ForceOperator
from pylabolt_b200.state import State, d2q9_constants, ghost_ring


def build_state(name):
    factory, kwargs, _ = cases.GOLDEN_CASES[name]
    sim = factory(**kwargs)
    return sim, State(sim, SingleComm(), 0, verbose=False)


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CASES))
def test_flags_and_fields_bit_exact(golden_dir, name):
    data = np.load(os.path.join(golden_dir, name + ".npz"))
    _, st = build_state(name)
    f = st.fields
    assert np.array_equal(st.domain.shape, data["shape"])
    for key in ("solid", "ghost_node", "solid_id", "solid_boundary",
                "fluid_boundary", "periodic_boundary"):
        assert np.array_equal(getattr(f, key), data[key]), key
    assert np.array_equal(f.surface_normals, data["surface_normals"])
    assert np.array_equal(f.density, data["density_0"])
    assert np.array_equal(f.velocity, data["velocity_0"])
    assert st.boundary.x_periodic == bool(data["x_periodic"])
    assert st.boundary.y_periodic == bool(data["y_periodic"])


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CASES))
def test_boundary_link_lists_bit_exact(golden_dir, name):
    data = np.load(os.path.join(golden_dir, name + ".npz"))
    _, st = build_state(name)
    elements = st.boundary.boundary_elements
    assert len(elements) == int(data["n_elements"])
    for n, el in enumerate(elements):
        assert el.name == str(data[f"el{n}_name"])
        assert el.type_fluid == str(data[f"el{n}_type"])
        assert el.boundary_nodes.dtype == np.int64
        assert np.array_equal(el.boundary_nodes, data[f"el{n}_nodes"])
        assert np.array_equal(el.out_list, data[f"el{n}_out"])
        assert np.array_equal(el.inv_list, data[f"el{n}_inv"])
        assert np.array_equal(el.normal, data[f"el{n}_normal"])
        assert np.array_equal(el.vector_fluid, data[f"el{n}_vector"])
        assert float(el.scalar_fluid) == float(data[f"el{n}_scalar"])


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CASES))
def test_collision_and_forcing_parameters(golden_dir, name):
    data = np.load(os.path.join(golden_dir, name + ".npz"))
    sim, st = build_state(name)
    col = CollisionOperator(sim, FluidLB(), st, SingleComm(), verbose=False)
    frc = ForceOperator(sim, FluidLB(), st, SingleComm(),
                        collision_operator=col, verbose=False)
    assert float(col.omega_fluid) == float(data["omega"])
    assert np.array_equal(frc.gravity, data["gravity"])


def test_lattice_constants(golden_dir):
    """tests/unit/test_lattice.py:94-135 of the reference + probed values."""
    data = np.load(os.path.join(golden_dir, "cavity.npz"))
    c = d2q9_constants()
    assert np.array_equal(
        [c["cs"], c["cs_2"], c["inv_cs_2"], c["inv_cs_4"]],
        data["lattice_consts"][:4])
    assert np.array_equal(c["weights"], data["weights"])
    assert c["cx"].tolist() == [0, 1, 0, -1, 0, 1, -1, -1, 1]
    assert c["cy"].tolist() == [0, 0, 1, 0, -1, 1, 1, -1, -1]
    assert c["inv_list"].tolist() == [0, 3, 4, 1, 2, 7, 8, 5, 6]
    for k in range(9):
        assert c["cx"][c["inv_list"][k]] == -c["cx"][k]
        assert c["cy"][c["inv_list"][k]] == -c["cy"][k]
    assert abs(c["weights"].sum() - 1.0) < 1e-15


def test_ghost_ring_matches_loop():
    """Fields.init_ghost_nodes, tests/unit/test_fields.py:61-86."""
    for shape in ((3, 3), (4, 7), (9, 5)):
        ghost = ghost_ring(shape)
        expect = np.zeros(shape[0] * shape[1], dtype=bool)
        for ind in range(expect.size):
            i, j = divmod(ind, shape[1])
            expect[ind] = i == 0 or j == 0 or i == shape[0] - 1 or \
                j == shape[1] - 1
        assert np.array_equal(ghost, expect)
