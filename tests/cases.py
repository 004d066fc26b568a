"""Case files (the reference's ``simulation.py`` schema) shared by the golden
generator (which feeds them to the reference itself) and by the parity tests
(which feed them to pylabolt_b200 and to the oracle).

Every function returns a ``types.SimpleNamespace`` carrying the same
module-level dicts a user's ``simulation.py`` defines (reference keywords:
SURVEY.md section 8(b), "Case-file keywords that must keep working").

Only plain python / math is used inside ``func`` initialisers so that the
reference (which calls them in a per-node python loop,
pylabolt/base/init_fields.py:336-346) and our vectorised initialiser see the
same values.
"""
import math
from types import SimpleNamespace


def _control(end_time):
    return {
        "start_time": 0,
        "end_time": end_time,
        "std_out_interval": None,
        "save_interval": None,
        "checkpoint_interval": None,
        "precision": "double",
    }


def _walls(nx, ny):
    return {
        "left": [[0, 0], [0, ny - 1]],
        "right": [[nx - 1, 0], [nx - 1, ny - 1]],
        "bottom": [[0, 0], [nx - 1, 0]],
        "top": [[0, ny - 1], [nx - 1, ny - 1]],
    }


def _perturbed_velocity(amp, nx, ny):
    def func(i, j):
        return (amp * math.sin(2.0 * math.pi * (i + 0.5) / nx) *
                math.cos(2.0 * math.pi * (j + 0.25) / ny),
                amp * math.cos(2.0 * math.pi * (i + 0.125) / nx) *
                math.sin(2.0 * math.pi * (j + 0.75) / ny))
    return func


def _perturbed_density(amp, nx, ny):
    def func(i, j):
        return 1.0 + amp * math.cos(2.0 * math.pi * i / nx) * \
            math.cos(4.0 * math.pi * j / ny)
    return func


def cavity(nx=33, ny=29, end_time=60, lid=0.1, kin_visc=0.1, model="BGK"):
    """Lid-driven cavity, docs/Setup.rst:15-64 (tutorials/cavity/Re_100 is
    101x101, nu=0.1): three bounce_back walls + moving lid."""
    seg = _walls(nx, ny)
    return SimpleNamespace(
        name=f"cavity_{nx}x{ny}_{model}",
        control_dict=_control(end_time),
        mesh_dict={"grid": [nx, ny]},
        lattice_dict={"lattice_type": "D2Q9"},
        decompose_dict={"nx": 1, "ny": 1},
        transport_dict={"kin_visc": kin_visc},
        initial_fields_dict={"default": {"fluid": {
            "velocity": {"type": "fixed", "value": [0.0, 0.0]},
            "density": {"type": "fixed", "value": 1.0},
            "pressure": {"type": "fixed", "value": 0.0},
        }}},
        boundary_dict={
            "options": {},
            "walls": {"wall": True,
                      "segments": [seg["left"], seg["right"], seg["bottom"]],
                      "fluid": {"type": "bounce_back"}},
            "lid": {"wall": True, "segments": [seg["top"]],
                    "fluid": {"type": "fixed_velocity", "value": [lid, 0.0]}},
        },
        obstacle_dict={"options": {}},
        collision_dict={"fluid": {"model": model,
                                  "equilibrium": "density_based_second_order",
                                  "forcing_model": None}},
        forcing_dict={},
    )


def poiseuille(nx=24, ny=21, end_time=60, forcing="guo_second_order",
               g=1.0e-5, kin_visc=0.064, model="BGK", perturb=0.02):
    """Gravity-driven plane Poiseuille flow: x-periodic pair, bounce_back
    top/bottom, Guo forcing.  A smooth perturbation makes the flow depend on
    x so that the periodic wrap is observable."""
    seg = _walls(nx, ny)
    return SimpleNamespace(
        name=f"poiseuille_{nx}x{ny}_{model}_{forcing}",
        control_dict=_control(end_time),
        mesh_dict={"grid": [nx, ny]},
        lattice_dict={"lattice_type": "D2Q9"},
        decompose_dict={"nx": 1, "ny": 1},
        transport_dict={"kin_visc": kin_visc},
        initial_fields_dict={"default": {"fluid": {
            "velocity": {"type": "func",
                         "func": _perturbed_velocity(perturb, nx, ny)},
            "density": {"type": "func",
                        "func": _perturbed_density(0.01, nx, ny)},
            "pressure": {"type": "fixed", "value": 0.0},
        }}},
        boundary_dict={
            "options": {},
            "inout": {"wall": False,
                      "segments": [seg["left"], seg["right"]],
                      "fluid": {"type": "periodic"}},
            "plates": {"wall": True,
                       "segments": [seg["bottom"], seg["top"]],
                       "fluid": {"type": "bounce_back"}},
        },
        obstacle_dict={"options": {}},
        collision_dict={"fluid": {"model": model,
                                  "equilibrium": "density_based_second_order",
                                  "forcing_model": forcing}},
        forcing_dict={"gravity": [g, 0.0]},
    )


def channel_y(nx=19, ny=26, end_time=50, forcing="guo_linear", g=2.0e-5):
    """Same physics turned by 90 degrees: y-periodic pair, walls left/right,
    body force along y.  Exercises the top/bottom wrap."""
    seg = _walls(nx, ny)
    return SimpleNamespace(
        name=f"channel_y_{nx}x{ny}_{forcing}",
        control_dict=_control(end_time),
        mesh_dict={"grid": [nx, ny]},
        lattice_dict={"lattice_type": "D2Q9"},
        decompose_dict={"nx": 1, "ny": 1},
        transport_dict={"kin_visc": 0.05},
        initial_fields_dict={"default": {"fluid": {
            "velocity": {"type": "func",
                         "func": _perturbed_velocity(0.015, nx, ny)},
            "density": {"type": "fixed", "value": 1.0},
            "pressure": {"type": "fixed", "value": 0.0},
        }}},
        boundary_dict={
            "options": {},
            "sides": {"wall": True,
                      "segments": [seg["left"], seg["right"]],
                      "fluid": {"type": "bounce_back"}},
            "updown": {"wall": False,
                       "segments": [seg["bottom"], seg["top"]],
                       "fluid": {"type": "periodic"}},
        },
        obstacle_dict={"options": {}},
        collision_dict={"fluid": {"model": "BGK",
                                  "equilibrium": "density_based_second_order",
                                  "forcing_model": forcing}},
        forcing_dict={"gravity": [0.0, g]},
    )


def cylinder(nx=60, ny=31, end_time=60, model="BGK", rho_in=1.003,
             rho_out=0.997, radius=4, spin=None):
    """Flow past a cylinder: fixed_pressure (anti-bounce-back) inlet and
    outlet, bounce_back plates, static circle.  ``spin`` makes the circle a
    prescribed-rotation body so that the moving-wall term of the in-kernel
    bounce back (cpu/streaming_kernels.py:40-47) is non-zero; rotation does
    not change the rasterised mask."""
    seg = _walls(nx, ny)
    body = {"type": "circle", "radius": radius,
            "center": [nx // 4, ny // 2], "density": 1.0, "static": True}
    options = {}
    if spin is not None:
        body["static"] = False
        body["solid_motion_dict"] = {
            "type": "fixed_velocity", "degree_of_freedom": "rotation",
            "linear_velocity": [0.0, 0.0], "angular_velocity": spin}
        options = {"compute_force_torque": True}
    return SimpleNamespace(
        name=f"cylinder_{nx}x{ny}_{model}" + ("_spin" if spin else ""),
        control_dict=_control(end_time),
        mesh_dict={"grid": [nx, ny]},
        lattice_dict={"lattice_type": "D2Q9"},
        decompose_dict={"nx": 1, "ny": 1},
        transport_dict={"kin_visc": 0.04},
        initial_fields_dict={"default": {"fluid": {
            "velocity": {"type": "fixed", "value": [0.02, 0.0]},
            "density": {"type": "fixed", "value": 1.0},
            "pressure": {"type": "fixed", "value": 0.0},
        }}},
        boundary_dict={
            "options": {},
            "inlet": {"wall": False, "segments": [seg["left"]],
                      "fluid": {"type": "fixed_pressure", "value": rho_in}},
            "outlet": {"wall": False, "segments": [seg["right"]],
                       "fluid": {"type": "fixed_pressure", "value": rho_out}},
            "plates": {"wall": True,
                       "segments": [seg["bottom"], seg["top"]],
                       "fluid": {"type": "bounce_back"}},
        },
        obstacle_dict={"options": options, "cyl": body},
        collision_dict={"fluid": {"model": model,
                                  "equilibrium": "density_based_second_order",
                                  "forcing_model": None}},
        forcing_dict={},
    )


def inflow_cylinder(nx=48, ny=27, end_time=50):
    """fixed_velocity inlet, fixed_pressure outlet, ellipse obstacle; walls
    listed first so the inlet/outlet win the corner links."""
    seg = _walls(nx, ny)
    return SimpleNamespace(
        name=f"inflow_ellipse_{nx}x{ny}",
        control_dict=_control(end_time),
        mesh_dict={"grid": [nx, ny]},
        lattice_dict={"lattice_type": "D2Q9"},
        decompose_dict={"nx": 1, "ny": 1},
        transport_dict={"kin_visc": 0.05},
        initial_fields_dict={"default": {"fluid": {
            "velocity": {"type": "fixed", "value": [0.03, 0.0]},
            "density": {"type": "fixed", "value": 1.0},
            "pressure": {"type": "fixed", "value": 0.0},
        }}},
        boundary_dict={
            "options": {},
            "plates": {"wall": True,
                       "segments": [seg["bottom"], seg["top"]],
                       "fluid": {"type": "bounce_back"}},
            "inlet": {"wall": False, "segments": [seg["left"]],
                      "fluid": {"type": "fixed_velocity",
                                "value": [0.03, 0.0]}},
            "outlet": {"wall": False, "segments": [seg["right"]],
                       "fluid": {"type": "fixed_pressure", "value": 1.0}},
        },
        obstacle_dict={"options": {}, "body": {
            "type": "ellipse", "semi_major_axis": 5.0,
            "semi_minor_axis": 3.0, "inclination_angle": 25.0,
            "center": [14, 13], "density": 1.0, "static": True}},
        collision_dict={"fluid": {"model": "BGK",
                                  "equilibrium": "density_based_second_order",
                                  "forcing_model": None}},
        forcing_dict={},
    )


def periodic_box(nx=28, ny=22, end_time=50, forcing="guo_second_order",
                 model="BGK"):
    """Doubly periodic box with a circle that straddles the x and the y wrap
    (minimum-image rasterisation, cpu/obstacle_kernels.py:13-40, and solid
    ghost flags, base/obstacle_operator.py:36-41)."""
    seg = _walls(nx, ny)
    return SimpleNamespace(
        name=f"periodic_box_{nx}x{ny}_{model}_{forcing}",
        control_dict=_control(end_time),
        mesh_dict={"grid": [nx, ny]},
        lattice_dict={"lattice_type": "D2Q9"},
        decompose_dict={"nx": 1, "ny": 1},
        transport_dict={"kin_visc": 0.08},
        initial_fields_dict={"default": {"fluid": {
            "velocity": {"type": "func",
                         "func": _perturbed_velocity(0.02, nx, ny)},
            "density": {"type": "func",
                        "func": _perturbed_density(0.005, nx, ny)},
            "pressure": {"type": "fixed", "value": 0.0},
        }}},
        boundary_dict={
            "options": {},
            "lr": {"wall": False, "segments": [seg["left"], seg["right"]],
                   "fluid": {"type": "periodic"}},
            "bt": {"wall": False, "segments": [seg["bottom"], seg["top"]],
                   "fluid": {"type": "periodic"}},
        },
        obstacle_dict={"options": {}, "ball": {
            "type": "circle", "radius": 4.5, "center": [1.0, 20.5],
            "density": 1.0, "static": True}},
        collision_dict={"fluid": {"model": model,
                                  "equilibrium": "density_based_second_order",
                                  "forcing_model": forcing}},
        forcing_dict={"gravity": [1.5e-5, -0.5e-5]},
    )


def open_corner(nx=17, ny=15, end_time=30):
    """Pathological on purpose: the bottom wall stops short of both corners
    and there is no right boundary at all, so some edge links are covered by
    no element and keep the 0.0 pulled from the ghost ring (SURVEY.md App. A,
    'Consequences worth knowing')."""
    seg = _walls(nx, ny)
    return SimpleNamespace(
        name=f"open_corner_{nx}x{ny}",
        control_dict=_control(end_time),
        mesh_dict={"grid": [nx, ny]},
        lattice_dict={"lattice_type": "D2Q9"},
        decompose_dict={"nx": 1, "ny": 1},
        transport_dict={"kin_visc": 0.1},
        initial_fields_dict={"default": {"fluid": {
            "velocity": {"type": "fixed", "value": [0.01, -0.005]},
            "density": {"type": "fixed", "value": 1.0},
            "pressure": {"type": "fixed", "value": 0.0},
        }}},
        boundary_dict={
            "options": {},
            "lid": {"wall": True, "segments": [seg["top"]],
                    "fluid": {"type": "fixed_velocity",
                              "value": [0.05, 0.0]}},
            "floor": {"wall": True,
                      "segments": [[[2, 0], [nx - 3, 0]]],
                      "fluid": {"type": "bounce_back"}},
            "left": {"wall": True, "segments": [seg["left"]],
                     "fluid": {"type": "bounce_back"}},
        },
        obstacle_dict={"options": {}},
        collision_dict={"fluid": {"model": "BGK",
                                  "equilibrium": "density_based_second_order",
                                  "forcing_model": None}},
        forcing_dict={},
    )


# name -> (factory, kwargs, steps at which the reference state is recorded)
GOLDEN_CASES = {
    "cavity": (cavity, {}, (1, 10, 60)),
    "poiseuille_guo2": (poiseuille, {"forcing": "guo_second_order"},
                        (1, 10, 60)),
    "poiseuille_guo1": (poiseuille, {"forcing": "guo_linear"}, (1, 10, 60)),
    "channel_y": (channel_y, {}, (1, 50)),
    "cylinder": (cylinder, {}, (1, 10, 60)),
    "cylinder_spin": (cylinder, {"spin": 0.01}, (1, 10, 60)),
    "inflow_ellipse": (inflow_cylinder, {}, (1, 50)),
    "periodic_box": (periodic_box, {}, (1, 10, 50)),
    "open_corner": (open_corner, {}, (1, 8)),
}
