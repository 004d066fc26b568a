"""Case-file validation: same keywords and the same error messages as the
reference's containers (modelled on the reference's tests/unit/test_control.py,
test_mesh.py, test_lattice.py, test_boundary.py, test_obstacle.py,
test_init_fields.py, which check ValueError messages)."""
import copy

import numpy as np
import pytest

import cases
from pylabolt_b200.comm import SingleComm
from pylabolt_b200.operators import CollisionOperator, FluidLB, ForceOperator
from pylabolt_b200.state import Control, Lattice, Mesh, State


def state_of(sim):
    return State(sim, SingleComm(), 0, verbose=False)


def broken(mutate, factory=cases.cavity):
    sim = factory()
    sim = copy.copy(sim)
    for name in ("control_dict", "mesh_dict", "lattice_dict", "boundary_dict",
                 "obstacle_dict", "initial_fields_dict", "collision_dict",
                 "forcing_dict", "transport_dict", "decompose_dict"):
        setattr(sim, name, copy.deepcopy(getattr(sim, name)))
    mutate(sim)
    return sim


@pytest.mark.parametrize("key", Control.KEYS)
def test_control_missing_key(key):
    sim = broken(lambda s: s.control_dict.pop(key))
    with pytest.raises(ValueError, match=key + " missing in control_dict"):
        Control(sim, 0, verbose=False)


def test_control_precision():
    sim = broken(lambda s: s.control_dict.update(precision="half"))
    with pytest.raises(ValueError, match="unsupported precision"):
        Control(sim, 0, verbose=False)
    # 'single' is a valid container value (reference: control.py:43-44); the
    # b200 State refuses to build the case
    sim = broken(lambda s: s.control_dict.update(precision="single"))
    assert Control(sim, 0, verbose=False).precision is np.float32
    with pytest.raises(ValueError, match="fp64"):
        state_of(sim)
    c = Control(cases.cavity(), 0, verbose=False)
    assert c.precision is np.float64
    assert c.float_min == np.finfo(np.float64).eps


@pytest.mark.parametrize("grid,message", [
    (None, "grid missing in mesh_dict"),
    ((4, 4), "must be a list"),
    ([4, 4, 4], "must be a list"),
    ([0, 4], "cannot be zero"),
    ([1, 1], "is a point"),
])
def test_mesh_validation(grid, message):
    def mutate(s):
        if grid is None:
            s.mesh_dict.pop("grid")
        else:
            s.mesh_dict["grid"] = grid
    with pytest.raises(ValueError, match=message):
        Mesh(broken(mutate), 0, verbose=False)


def test_mesh_dimensions():
    assert Mesh(cases.cavity(), 0, verbose=False).dimensions == 2
    sim = broken(lambda s: s.mesh_dict.update(grid=[1, 9]))
    assert Mesh(sim, 0, verbose=False).dimensions == 1


def test_lattice_validation():
    sim = cases.cavity()
    control, mesh = Control(sim, 0, False), Mesh(sim, 0, False)
    with pytest.raises(ValueError, match="lattice_type missing"):
        Lattice(broken(lambda s: s.lattice_dict.clear()), control, mesh, 0, False)
    with pytest.raises(ValueError, match="Unsupported lattice type"):
        Lattice(broken(lambda s: s.lattice_dict.update(lattice_type="D3Q19")),
                control, mesh, 0, False)
    line = Mesh(broken(lambda s: s.mesh_dict.update(grid=[1, 9])), 0, False)
    with pytest.raises(ValueError, match="incompatible"):
        Lattice(sim, control, line, 0, False)
    # the container knows D1Q3 (reference: lattice.py:61-70) ...
    d1q3 = broken(lambda s: s.lattice_dict.update(lattice_type="D1Q3"))
    with pytest.raises(ValueError, match="incompatible"):
        Lattice(d1q3, control, mesh, 0, False)
    lat = Lattice(d1q3, control, line, 0, False)
    assert lat.no_of_directions == 3 and list(lat.inv_list) == [0, 2, 1]
    # ... and the b200 State refuses to build such a case
    d1q3.mesh_dict.update(grid=[9, 1])
    with pytest.raises(ValueError, match="D1Q3 is not available"):
        state_of(d1q3)


@pytest.mark.parametrize("mutate,message", [
    (lambda s: delattr(s, "boundary_dict"), "boundary_dict not found"),
    (lambda s: s.boundary_dict.pop("options"), "options missing"),
    (lambda s: s.boundary_dict["lid"].pop("wall"), "wall missing in boundary: lid"),
    (lambda s: s.boundary_dict["lid"].update(wall="yes"), "wall entry must be True/False"),
    (lambda s: s.boundary_dict["lid"].pop("segments"), "segments missing"),
    (lambda s: s.boundary_dict["lid"].update(segments=[]), "cannot be an empty list"),
    (lambda s: s.boundary_dict["lid"].update(segments=[[0, 1]]), "must have structure"),
    (lambda s: s.boundary_dict["lid"].update(segments=[[[5, 28], [1, 28]]]),
     "x2 >= x1"),
    (lambda s: s.boundary_dict["lid"].update(segments=[[[0, 0], [5, 5]]]),
     "axis-aligned"),
    (lambda s: s.boundary_dict["lid"].update(segments=[[[3, 5], [9, 5]]]),
     "domain edge"),
    (lambda s: s.boundary_dict["lid"].pop("fluid"), "fluid missing in boundary"),
    (lambda s: s.boundary_dict["lid"]["fluid"].pop("type"), "type missing"),
    (lambda s: s.boundary_dict["lid"]["fluid"].update(type="slip"),
     "Unsupported boundary condition for fluid: slip"),
    (lambda s: s.boundary_dict["lid"]["fluid"].pop("value"), "value missing"),
    (lambda s: s.boundary_dict["lid"]["fluid"].update(value=0.1),
     "must be a list"),
    (lambda s: s.boundary_dict["options"].update(compute_force=1),
     "compute_force must be a bool"),
    (lambda s: s.boundary_dict["options"].update(write_boundary_data={}),
     "interval missing"),
    (lambda s: s.boundary_dict["options"].update(
        write_boundary_data={"interval": 0}), "must be > 0"),
])
def test_boundary_validation(mutate, message):
    with pytest.raises(ValueError, match=message):
        state_of(broken(mutate))


def test_fixed_pressure_value_type():
    def mutate(s):
        s.boundary_dict["inlet"]["fluid"]["value"] = [1.0]
    with pytest.raises(ValueError, match="float or int"):
        state_of(broken(mutate, cases.cylinder))


@pytest.mark.parametrize("mutate,message", [
    (lambda s: s.boundary_dict["inout"].update(
        segments=[s.boundary_dict["inout"]["segments"][0]]), "periodic pair"),
    (lambda s: s.boundary_dict["inout"].update(
        segments=[[[0, 0], [0, 20]], [[0, 20], [23, 20]]]),
     "different orientation"),
    (lambda s: s.boundary_dict["inout"].update(
        segments=[[[0, 0], [0, 10]], [[23, 0], [23, 10]]]),
     "entire y-direction"),
    (lambda s: s.boundary_dict["inout"].update(
        segments=[[[0, 0], [0, 20]], [[0, 0], [0, 20]]]),
     "left-right"),
])
def test_periodic_pair_validation(mutate, message):
    with pytest.raises(ValueError, match=message):
        state_of(broken(mutate, cases.poiseuille))


def test_legacy_boundary_spellings_are_aliases():
    def mutate(s):
        s.boundary_dict["walls"]["fluid"]["type"] = "bounceBack"
        s.boundary_dict["lid"]["fluid"]["type"] = "fixedU"
    st = state_of(broken(mutate))
    kinds = [el.type_fluid for el in st.boundary.boundary_elements]
    assert kinds == ["bounce_back"] * 3 + ["fixed_velocity"]


@pytest.mark.parametrize("mutate,message", [
    (lambda s: s.obstacle_dict.pop("options"), "options missing in obstacle_dict"),
    (lambda s: s.obstacle_dict["cyl"].pop("type"), "type missing in obstacle"),
    (lambda s: s.obstacle_dict["cyl"].update(type="square"),
     "Unsupported obstacle type: square"),
    (lambda s: s.obstacle_dict["cyl"].pop("radius"), "radius missing"),
    (lambda s: s.obstacle_dict["cyl"].update(radius="4"), "radius must be float or int"),
    (lambda s: s.obstacle_dict["cyl"].pop("center"), "center missing"),
    (lambda s: s.obstacle_dict["cyl"].update(center=(3, 4)), "center must be list"),
    (lambda s: s.obstacle_dict["cyl"].pop("density"), "density missing"),
    (lambda s: s.obstacle_dict["cyl"].pop("static"), "static missing"),
    (lambda s: s.obstacle_dict["cyl"].update(static=0), "static must be True/False"),
    (lambda s: s.obstacle_dict["cyl"].update(static=False),
     "solid_motion_dict missing"),
    (lambda s: s.obstacle_dict["options"].update(compute_force_torque="y"),
     "compute_force_torque must be bool"),
])
def test_obstacle_validation(mutate, message):
    with pytest.raises(ValueError, match=message):
        state_of(broken(mutate, cases.cylinder))


def test_moving_bodies_are_refused():
    def mutate(s):
        s.obstacle_dict["options"]["compute_force_torque"] = True
        s.obstacle_dict["cyl"]["static"] = False
        s.obstacle_dict["cyl"]["solid_motion_dict"] = {
            "type": "calculated", "degree_of_freedom": "both",
            "linear_velocity": [0.0, 0.0], "angular_velocity": 0.0}
    with pytest.raises(ValueError, match="moving obstacles are not supported"):
        state_of(broken(mutate, cases.cylinder))


def test_non_static_needs_force_torque():
    sim = cases.cylinder(spin=0.01)
    sim.obstacle_dict["options"] = {}
    with pytest.raises(ValueError, match="compute_force_torque must be True"):
        state_of(sim)


@pytest.mark.parametrize("mutate,message", [
    (lambda s: s.initial_fields_dict.pop("default"), "default missing"),
    (lambda s: s.initial_fields_dict["default"].pop("fluid"), "fluid missing"),
    (lambda s: s.initial_fields_dict["default"]["fluid"].pop("density"),
     "density"),
    (lambda s: s.initial_fields_dict["default"]["fluid"]["density"].pop("type"),
     "type missing in field definition"),
    (lambda s: s.initial_fields_dict["default"]["fluid"]["density"].update(
        type="file"), "Unsupported velocity initialization"),
    (lambda s: s.initial_fields_dict["default"]["fluid"]["density"].pop("value"),
     "value missing for fixed type"),
    (lambda s: s.initial_fields_dict["default"]["fluid"]["velocity"].update(
        value=0.0), "vector value must be a list"),
    (lambda s: s.initial_fields_dict["default"]["fluid"]["velocity"].update(
        type="func"), "func missing"),
])
def test_initial_fields_validation(mutate, message):
    with pytest.raises(ValueError, match=message):
        state_of(broken(mutate))


def test_region_override_applies_after_default():
    def mutate(s):
        s.initial_fields_dict["jet"] = {"fluid": {"density": {
            "type": "func", "func": lambda i, j: 1.5 if i < 3 else 1.0}}}
    st = state_of(broken(mutate))
    rho = st.fields.density.reshape(tuple(st.domain.shape))
    assert np.all(rho[1:4, 1:-1] == 1.5) and np.all(rho[4:-1, 1:-1] == 1.0)
    assert np.all(rho[0] == 0.0)


@pytest.mark.parametrize("mutate,message", [
    (lambda s: delattr(s, "collision_dict"), "collision_dict not found"),
    (lambda s: s.collision_dict.pop("fluid"), "fluid missing in collision_dict"),
    (lambda s: s.collision_dict["fluid"].pop("model"), "model missing"),
    (lambda s: s.collision_dict["fluid"].update(model="TRT"),
     "Unsupported fluid collision model: TRT"),
    (lambda s: s.collision_dict["fluid"].update(equilibrium="incompressible"),
     "Unsupported fluid equilibrium model"),
    (lambda s: s.collision_dict["fluid"].update(forcing_model="shan_chen"),
     "Unsupported fluid forcing model"),
])
def test_collision_validation(mutate, message):
    sim = broken(mutate)
    st = state_of(cases.cavity())
    with pytest.raises(ValueError, match=message):
        CollisionOperator(sim, FluidLB(), st, SingleComm(), verbose=False)


def test_gravity_is_ignored_without_forcing_model(capsys):
    sim = cases.cavity()
    sim.forcing_dict = {"gravity": [1e-5, 0.0]}
    st = state_of(sim)
    col = CollisionOperator(sim, FluidLB(), st, SingleComm(), verbose=False)
    frc = ForceOperator(sim, FluidLB(), st, SingleComm(),
                        collision_operator=col, verbose=True)
    assert np.array_equal(frc.gravity, [0.0, 0.0])
    assert "gravity ignored" in capsys.readouterr().out


def test_tau_and_omega_follow_the_reference_formula():
    sim = cases.cavity()
    st = state_of(sim)
    col = CollisionOperator(sim, FluidLB(), st, SingleComm(), verbose=False)
    assert col.tau_fluid == np.float64(0.1) * st.lattice.inv_cs_2 + 0.5
    assert col.omega_fluid == 1.25
    assert col.tau_fluid == 0.7999999999999999


def test_transport_validation():
    with pytest.raises(ValueError, match="kin_visc missing"):
        state_of(broken(lambda s: s.transport_dict.clear()))
    with pytest.raises(ValueError, match="kin_visc must be a float/int"):
        state_of(broken(lambda s: s.transport_dict.update(kin_visc="0.1")))
