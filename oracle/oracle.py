"""ctypes front end of the CPU oracle (oracle/plb_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of plb_oracle.c.  Imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by pylabolt_b200.

The arrays live in the reference's own layouts (pylabolt/base/fields.py:50-92):
flat index ind = x * Ny_pad + y, populations (size, 9), vectors (size, 2).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

BC_TYPES = {"bounce_back": 0, "fixed_velocity": 1, "fixed_pressure": 2,
            "periodic": 3, "zero_gradient": 4}
FORCING = {None: 0, "None": 0, "guo_linear": 1, "guo_second_order": 2}
COLLISION = {"BGK": 0, "MRT": 1}

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_u8_p = ctypes.POINTER(ctypes.c_uint8)
_c_i64_p = ctypes.POINTER(ctypes.c_int64)


class _Params(ctypes.Structure):
    _fields_ = [
        ("nx_pad", ctypes.c_int64), ("ny_pad", ctypes.c_int64),
        ("inv_cs_2", ctypes.c_double), ("inv_cs_4", ctypes.c_double),
        ("float_min", ctypes.c_double),
        ("weights", ctypes.c_double * 9),
        ("omega", ctypes.c_double),
        ("gravity", ctypes.c_double * 2),
        ("s", ctypes.c_double * 9),
        ("collision", ctypes.c_int32), ("forcing", ctypes.c_int32),
        ("x_periodic", ctypes.c_int32), ("y_periodic", ctypes.c_int32),
    ]


class _Fields(ctypes.Structure):
    _fields_ = [
        ("solid", _c_u8_p), ("ghost", _c_u8_p),
        ("density", _c_double_p), ("velocity", _c_double_p),
        ("force", _c_double_p), ("pop", _c_double_p),
        ("pop_new", _c_double_p),
    ]


class _Element(ctypes.Structure):
    _fields_ = [
        ("type", ctypes.c_int32),
        ("n_nodes", ctypes.c_int64),
        ("nodes", _c_i64_p),
        ("out_list", ctypes.c_int64 * 3), ("inv_list", ctypes.c_int64 * 3),
        ("normal", ctypes.c_int64 * 2),
        ("vector", ctypes.c_double * 2),
        ("scalar", ctypes.c_double),
    ]


def build(force=False):
    """Compile liboracle.so with oracle/Makefile (gcc, OpenMP, no FMA)."""
    deps = [os.path.join(_HERE, "plb_oracle.c"), os.path.join(_HERE, "Makefile")]
    if (force or not os.path.exists(_LIB_PATH) or
            any(os.path.getmtime(_LIB_PATH) < os.path.getmtime(d) for d in deps)):
        subprocess.check_call(["make", "-B", "-C", _HERE, "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def lattice_constants():
    """base/lattice.py:41-60 and base/control.py:49, computed the same way."""
    cs = np.float64(1 / np.sqrt(3))
    cs_2 = cs * cs
    inv_cs_2 = 1.0 / cs_2
    inv_cs_4 = inv_cs_2 * inv_cs_2
    weights = np.array([4 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 9,
                        1 / 36, 1 / 36, 1 / 36, 1 / 36], dtype=np.float64)
    return {"cs": cs, "cs_2": cs_2, "inv_cs_2": inv_cs_2,
            "inv_cs_4": inv_cs_4, "weights": weights,
            "float_min": np.finfo(np.float64).eps}


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


class Oracle:
    """One rank's reference-layout state plus the reference step on it.

    Parameters mirror what the reference's operators read from ``State``:
    shape (domain.shape), solid / ghost_node flags, density and velocity
    after init_fields + obstacle rasterisation, the boundary elements in
    boundary_dict order, omega, gravity, the forcing / collision names.
    """

    def __init__(self, shape, solid, ghost_node, density, velocity, elements,
                 omega, gravity=(0.0, 0.0), forcing=None, collision="BGK",
                 x_periodic=False, y_periodic=False, mrt_rates=None,
                 n_threads=None):
        self.lib = lib()
        if n_threads:
            self.lib.oracle_set_threads(int(n_threads))
        consts = lattice_constants()
        self.nx_pad, self.ny_pad = int(shape[0]), int(shape[1])
        size = self.nx_pad * self.ny_pad
        self.size = size
        self.solid = np.ascontiguousarray(solid, dtype=np.uint8).copy()
        self.ghost = np.ascontiguousarray(ghost_node, dtype=np.uint8).copy()
        self.density = np.ascontiguousarray(density, dtype=np.float64).copy()
        self.velocity = np.ascontiguousarray(
            velocity, dtype=np.float64).reshape(size, 2).copy()
        self.force = np.zeros((size, 2), dtype=np.float64)
        self.pop = np.zeros((size, 9), dtype=np.float64)
        self.pop_new = np.zeros((size, 9), dtype=np.float64)
        self.density_old = np.zeros(size, dtype=np.float64)
        self.velocity_old = np.zeros((size, 2), dtype=np.float64)

        p = _Params()
        p.nx_pad, p.ny_pad = self.nx_pad, self.ny_pad
        p.inv_cs_2, p.inv_cs_4 = consts["inv_cs_2"], consts["inv_cs_4"]
        p.float_min = consts["float_min"]
        p.weights[:] = consts["weights"].tolist()
        p.omega = float(omega)
        p.gravity[:] = [float(gravity[0]), float(gravity[1])]
        if mrt_rates is None:
            # base/collision_operator.py:159-163
            mrt_rates = [1.0] * 7 + [float(omega)] * 2
        p.s[:] = [float(v) for v in mrt_rates]
        p.collision = COLLISION[collision]
        p.forcing = FORCING[forcing]
        p.x_periodic, p.y_periodic = int(bool(x_periodic)), int(bool(y_periodic))
        self.params = p

        self.fields = _Fields(
            _ptr(self.solid, _c_u8_p), _ptr(self.ghost, _c_u8_p),
            _ptr(self.density, _c_double_p), _ptr(self.velocity, _c_double_p),
            _ptr(self.force, _c_double_p), _ptr(self.pop, _c_double_p),
            _ptr(self.pop_new, _c_double_p))

        self._keep = []
        arr = (_Element * max(1, len(elements)))()
        for n, el in enumerate(elements):
            nodes = np.ascontiguousarray(el["nodes"], dtype=np.int64)
            self._keep.append(nodes)
            arr[n].type = BC_TYPES[el["type"]]
            arr[n].n_nodes = nodes.shape[0]
            arr[n].nodes = _ptr(nodes, _c_i64_p)
            arr[n].out_list[:] = [int(v) for v in el["out"]]
            arr[n].inv_list[:] = [int(v) for v in el["inv"]]
            arr[n].normal[:] = [int(v) for v in el["normal"]]
            arr[n].vector[:] = [float(v) for v in el.get("vector", (0, 0))]
            arr[n].scalar = float(el.get("scalar", 0.0))
        self.elements = arr
        self.n_elements = len(elements)

    # -- the reference's phases ------------------------------------------
    def initialize_pop(self):
        self.lib.oracle_initialize_pop(ctypes.byref(self.params),
                                       ctypes.byref(self.fields))

    def step(self, n_steps=1):
        self.lib.oracle_step(ctypes.byref(self.params),
                             ctypes.byref(self.fields), self.elements,
                             ctypes.c_int32(self.n_elements),
                             ctypes.c_int64(n_steps))

    def residue_sums(self):
        out = (ctypes.c_double * 6)()
        self.lib.oracle_residue_sums(
            ctypes.byref(self.params), ctypes.byref(self.fields),
            _ptr(self.density_old, _c_double_p),
            _ptr(self.velocity_old, _c_double_p), out)
        return np.array(out[:], dtype=np.float64)

    def residues(self):
        """utils/residues.py:199-222 -> (res_density, res_ux, res_uy)."""
        s = self.last_residue_sums = self.residue_sums()
        eps = self.params.float_min
        return (np.sqrt(s[0] / (s[1] + eps)), np.sqrt(s[2] / (s[3] + eps)),
                np.sqrt(s[4] / (s[5] + eps)))

    def boundary_force(self, n):
        """Force on boundary element n (cpu/force_torque_kernels.py:100-141)."""
        out = (ctypes.c_double * 2)()
        self.lib.oracle_boundary_force(ctypes.byref(self.fields),
                                       ctypes.byref(self.elements[n]), out)
        return np.array(out[:])

    def obstacle_force_torque(self, solid_id, fluid_boundary, offset,
                              grid_global_shape, ref_point, current_solid_id):
        """(fx, fy, torque) on one obstacle (force_torque_kernels.py:11-94)."""
        sid = np.ascontiguousarray(solid_id, dtype=np.int64)
        fb = np.ascontiguousarray(fluid_boundary, dtype=np.uint8)
        off = (ctypes.c_int64 * 2)(int(offset[0]), int(offset[1]))
        grid = (ctypes.c_int64 * 2)(int(grid_global_shape[0]),
                                    int(grid_global_shape[1]))
        ref = (ctypes.c_double * 2)(float(ref_point[0]), float(ref_point[1]))
        out = (ctypes.c_double * 3)()
        self.lib.oracle_obstacle_force_torque(
            ctypes.byref(self.params), ctypes.byref(self.fields),
            _ptr(sid, _c_i64_p), _ptr(fb, _c_u8_p), off, grid, ref,
            ctypes.c_int64(int(current_solid_id)), out)
        return np.array(out[:])

    @property
    def max_threads(self):
        return int(self.lib.oracle_max_threads())


def elements_from_golden(data):
    """Boundary elements as stored by tests/golden/make_golden.py."""
    out = []
    for n in range(int(data["n_elements"])):
        out.append({
            "name": str(data[f"el{n}_name"]),
            "type": str(data[f"el{n}_type"]),
            "nodes": data[f"el{n}_nodes"],
            "out": data[f"el{n}_out"],
            "inv": data[f"el{n}_inv"],
            "normal": data[f"el{n}_normal"],
            "vector": data[f"el{n}_vector"],
            "scalar": float(data[f"el{n}_scalar"]),
        })
    return out
