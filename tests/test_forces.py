"""Momentum-exchange wall / obstacle forces.

CPU part: the oracle's restatement of cpu/force_torque_kernels.py and the
host-side reduction of pylabolt_b200.force_torque against values computed by
the reference's own kernels (tests/golden, wall_force_* / body_force_*).  The
reference's prange reductions have unspecified order, so 1e-12, not bit-exact.
"""
import os

import numpy as np
import pytest

import cases
from oracle.oracle import Oracle, elements_from_golden
from pylabolt_b200.comm import SingleComm
from pylabolt_b200.force_torque import MomentumExchange, _INV
from pylabolt_b200.state import State

FORCE_CASES = ["cavity", "cylinder", "cylinder_spin", "inflow_ellipse",
               "periodic_box", "poiseuille_guo2"]


def close(a, b, scale=None):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.abs(b).max(), 1e-30) if scale is None else scale
    return np.abs(a - b).max() <= 1e-12 * scale + 1e-18


def setup(golden_dir, name):
    factory, kwargs, record = cases.GOLDEN_CASES[name]
    data = np.load(os.path.join(golden_dir, name + ".npz"))
    sim = factory(**kwargs)
    fluid = sim.collision_dict["fluid"]
    orc = Oracle(data["shape"], data["solid"], data["ghost_node"],
                 data["density_0"], data["velocity_0"],
                 elements_from_golden(data), float(data["omega"]),
                 gravity=data["gravity"], forcing=fluid["forcing_model"],
                 collision=fluid["model"], x_periodic=bool(data["x_periodic"]),
                 y_periodic=bool(data["y_periodic"]), n_threads=2)
    orc.initialize_pop()
    st = State(sim, SingleComm(), 0, verbose=False)
    return data, record, orc, st


@pytest.mark.parametrize("name", FORCE_CASES)
def test_oracle_forces_match_reference_kernels(golden_dir, name):
    data, record, orc, st = setup(golden_dir, name)
    done = 0
    for step in (0,) + tuple(record):
        orc.step(step - done)
        done = step
        want_wall = data[f"wall_force_{step}"]
        got_wall = np.array([orc.boundary_force(n)
                             for n in range(orc.n_elements)]).reshape(-1, 2)
        scale = max(np.abs(want_wall).max(), 1.0)
        assert close(got_wall, want_wall, scale), step
        want_body = data[f"body_force_{step}"]
        for n, body in enumerate(st.obstacle.obstacles):
            got = orc.obstacle_force_torque(
                st.fields.solid_id, st.fields.fluid_boundary, st.domain.offset,
                st.mesh.grid_global_shape, body.ref_point, body.id)
            assert close(got, want_body[n], max(np.abs(want_body).max(), 1.0))


@pytest.mark.parametrize("name", FORCE_CASES)
def test_host_reduction_matches_reference_kernels(golden_dir, name):
    """MomentumExchange fed with pop + pop_new pairs taken from the oracle's
    arrays (what the link kernel records on the device)."""
    data, record, orc, st = setup(golden_dir, name)
    fluid = np.flatnonzero(~st.fields.solid & ~st.fields.ghost_node)
    mom = MomentumExchange(st, fluid, every_element=True)   # kernel-level parity

    def exchange():
        return orc.pop[fluid][:, 1:] + orc.pop_new[fluid][:, _INV[1:]]

    assert np.array_equal(mom.initial_exchange(st.lattice), exchange())
    done = 0
    for step in (0,) + tuple(record):
        orc.step(step - done)
        done = step
        ex = exchange()
        want_wall = data[f"wall_force_{step}"]
        assert close(mom.boundary_forces(ex), want_wall,
                     max(np.abs(want_wall).max(), 1.0)), step
        want_body = data[f"body_force_{step}"]
        if want_body.size:
            assert close(mom.obstacle_forces(ex), want_body,
                         max(np.abs(want_body).max(), 1.0)), step


def test_history_files_have_the_reference_format(tmp_path):
    """utils/io_operator.py:250-357: header lines and 30-wide %.16e columns."""
    from pylabolt_b200.io_operator import InputOutputOperator
    from pylabolt_b200.operators import FluidLB
    sim = cases.cylinder(spin=0.01)
    sim.obstacle_dict["options"]["write_obstacle_data"] = {"interval": 2}
    sim.boundary_dict["options"] = {"compute_force": True,
                                    "write_boundary_data": {"interval": 1}}
    st = State(sim, SingleComm(), 0, verbose=False)
    io = InputOutputOperator(FluidLB(), st, None, SingleComm(), verbose=False,
                             root_dir=str(tmp_path))
    st.obstacle.obstacles[0].force[:] = [1.5, -2.0]
    st.obstacle.obstacles[0].torque = 0.25
    for t in range(0, 3):
        io.write_histories(st, t)
    body = (tmp_path / "output" / "histories" / "cyl.dat").read_text().splitlines()
    assert body[0].startswith("#     PyLaBolt obstacle history")
    assert body[5].split()[1:] == ["time", "pos_x", "pos_y", "alpha", "vel_x",
                                   "vel_y", "omega", "force_x", "force_y",
                                   "torque"]
    rows = body[6:]
    assert [int(r.split()[0]) for r in rows] == [0, 2]
    assert len(rows[0]) == 24 + 9 * 30
    assert float(rows[0].split()[7]) == 1.5 and float(rows[0].split()[9]) == 0.25
    walls = sorted(p.name for p in (tmp_path / "output" / "histories").iterdir())
    assert walls == ["cyl.dat", "plates_0.dat", "plates_1.dat"]   # wall=True only
    plate = (tmp_path / "output" / "histories" / "plates_0.dat").read_text()
    assert len(plate.splitlines()) == 4 + 3
