"""Host-side containers of a fluidLB case: the b200 mirror of the reference's
``State`` (pylabolt/base/state.py:14-147) and of the containers it builds.

Same case-file keywords, same attribute names, same validation messages as the
reference; the differences are deliberate and local:

* everything that the reference does with a per-node python loop
  (Fields.init_ghost_nodes, pylabolt/base/fields.py:166-179; ``func``
  initialisers, pylabolt/base/init_fields.py:325-375) is vectorised here, with
  bit-identical results;
* the population arrays and the force field are never allocated on the host --
  they live in HBM behind libplb;
* fp64 only (the north star is an fp64 path).
"""
import os

import numpy as np
from types import SimpleNamespace

from .helpers import print_log


def d2q9_constants(precision=np.float64):
    """pylabolt/base/lattice.py:41-60, computed the same way (so inv_cs_2 is
    2.999999999999999, not 3)."""
    cs = precision(1 / np.sqrt(3))
    cs_2 = cs * cs
    inv_cs_2 = 1.0 / cs_2
    return {
        "cs": cs, "cs_2": cs_2, "inv_cs_2": inv_cs_2,
        "inv_cs_4": inv_cs_2 * inv_cs_2,
        "cx": np.array([0, 1, 0, -1, 0, 1, -1, -1, 1], dtype=int),
        "cy": np.array([0, 0, 1, 0, -1, 1, 1, -1, -1], dtype=int),
        "weights": np.array([4 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 9,
                             1 / 36, 1 / 36, 1 / 36, 1 / 36], dtype=precision),
        "inv_list": np.array([0, 3, 4, 1, 2, 7, 8, 5, 6], dtype=int),
    }


def _require(module, name):
    if not hasattr(module, name):
        raise ValueError(name + " not found in simulation.py file")
    return getattr(module, name)


class Control:
    """control_dict -- pylabolt/base/control.py:6-54."""

    KEYS = ("start_time", "end_time", "std_out_interval", "save_interval",
            "checkpoint_interval", "precision")

    def __init__(self, simulation, rank=0, verbose=True):
        control_dict = _require(simulation, "control_dict")
        print_log("-" * 80, rank, verbose)
        print_log("Setting control parameters...", rank, verbose)
        for key in self.KEYS:
            if key not in control_dict:
                raise ValueError(key + " missing in control_dict")
        self.start_time = control_dict["start_time"]
        self.end_time = control_dict["end_time"]
        self.std_out_interval = control_dict["std_out_interval"]
        self.save_interval = control_dict["save_interval"]
        self.checkpoint_interval = control_dict["checkpoint_interval"]
        self.precision_type = control_dict["precision"]
        if self.precision_type == "double":
            self.precision = np.float64
        elif self.precision_type == "single":
            # a valid container value like in the reference; State refuses to
            # build an fp32 case for the b200 path (fp64 only)
            self.precision = np.float32
        else:
            raise ValueError("unsupported precision specified." +
                             "available precision (single, double)")
        self.float_min = np.finfo(self.precision).eps
        print_log("Setting control parameters done!", rank, verbose)
        print_log("-" * 80, rank, verbose)


class Mesh:
    """mesh_dict -- pylabolt/base/mesh.py:6-47."""

    def __init__(self, simulation, rank=0, verbose=True):
        mesh_dict = _require(simulation, "mesh_dict")
        if "grid" not in mesh_dict:
            raise ValueError("grid missing in mesh_dict")
        grid = mesh_dict["grid"]
        if not isinstance(grid, list) or len(grid) != 2:
            raise ValueError("grid entry in mesh_dict must be a list [Nx, Ny]")
        grid = np.array(grid, dtype=int)
        if np.any(grid == 0):
            raise ValueError("grid dimensions cannot be zero")
        if np.all(grid == 1):
            raise ValueError("grid specified is a point")
        self.dimensions = 1 if (grid[0] == 1 or grid[1] == 1) else 2
        self.grid_global_shape = grid
        self.grid_global_size = np.prod(grid)
        print_log("global grid size set: (" + str(grid) + ")", rank, verbose)


class Lattice:
    """lattice_dict -- pylabolt/base/lattice.py:6-75 (D2Q9 only here)."""

    def __init__(self, simulation, control, mesh, rank=0, verbose=True):
        lattice_dict = _require(simulation, "lattice_dict")
        if "lattice_type" not in lattice_dict:
            raise ValueError("lattice_type missing in lattice_dict")
        self.lattice_type = lattice_dict["lattice_type"]
        for key, value in d2q9_constants(control.precision).items():
            setattr(self, key, value)
        if self.lattice_type == "D2Q9":
            if mesh.dimensions != 2:
                raise ValueError(
                    "grid dimensions and lattice type are incompatible")
            self.no_of_directions = int(9)
        elif self.lattice_type == "D1Q3":
            # the container knows the reference's second lattice (:61-70);
            # State refuses to build a D1Q3 case for the b200 path
            if mesh.dimensions != 1:
                raise ValueError(
                    "grid dimensions and lattice type are incompatible")
            self.cx = np.array([0, 1, -1], dtype=int)
            self.cy = np.array([0, 0, 0], dtype=int)
            self.weights = np.array([2 / 3, 1 / 6, 1 / 6], dtype=control.precision)
            self.inv_list = np.array([0, 2, 1], dtype=int)
            self.no_of_directions = int(3)
        else:
            raise ValueError("Unsupported lattice type")
        print_log("lattice type set: " + self.lattice_type, rank, verbose)


class Domain:
    """decompose_dict -- pylabolt/parallel/domain.py:4-85.  Ranks are laid out
    i_proc = rank // ny, j_proc = rank % ny; every rank but the last gets
    ceil(N / n) nodes per direction and the last one the remainder."""

    def __init__(self, simulation, mesh, comm, verbose=True):
        self.mpi_rank = comm.Get_rank()
        self.mpi_size = comm.Get_size()
        decompose_dict = _require(simulation, "decompose_dict")
        if "nx" not in decompose_dict or "ny" not in decompose_dict:
            raise ValueError("nx or ny missing decompose_dict")
        self.no_of_procs_x = decompose_dict["nx"]
        self.no_of_procs_y = decompose_dict["ny"]
        if self.mpi_size != self.no_of_procs_x * self.no_of_procs_y:
            raise ValueError("invalid domain decomposition. " +
                             "nx * ny not equal to total no.of MPI processes")
        self.i_proc = self.mpi_rank // self.no_of_procs_y
        self.j_proc = self.mpi_rank % self.no_of_procs_y
        extent, offset = [], []
        for n_global, n_procs, i in (
                (int(mesh.grid_global_shape[0]), self.no_of_procs_x, self.i_proc),
                (int(mesh.grid_global_shape[1]), self.no_of_procs_y, self.j_proc)):
            chunk = int(np.ceil(n_global / n_procs))
            offset.append(i * chunk)
            extent.append(chunk if i != n_procs - 1 else n_global - i * chunk)
        self.Nx_rank, self.Ny_rank = extent
        self.offset = np.array(offset, dtype=int)
        self.Nx_pad = self.Nx_rank + 2
        self.Ny_pad = self.Ny_rank + 2
        self.shape = np.array([self.Nx_pad, self.Ny_pad])
        self.size = np.prod(self.shape)
        self.inner_shape = np.array([self.Nx_rank, self.Ny_rank])
        self.inner_size = np.prod(self.inner_shape)

    def require_slabs(self):
        """The b200 back end shards the lattice in x-slabs (y is the
        contiguous axis, so a face is three contiguous runs)."""
        if self.no_of_procs_y != 1:
            raise ValueError("the b200 back end decomposes the lattice in "
                             "x-slabs: decompose_dict must have ny = 1")
        if self.Nx_rank < 1:
            raise ValueError("invalid domain decomposition. a rank owns no "
                             "lattice columns")


class Transport:
    """transport_dict -- pylabolt/base/transport.py:4-86 (single-phase)."""

    def __init__(self, simulation, control, domain, verbose=True):
        rank = domain.mpi_rank
        print_log("-" * 80, rank, verbose)
        print_log("Setting transport properties...\n", rank, verbose)
        transport_dict = _require(simulation, "transport_dict")
        if "kin_visc" not in transport_dict:
            raise ValueError("kin_visc missing in transport_dict")
        kin_visc = transport_dict["kin_visc"]
        if not isinstance(kin_visc, (float, int)):
            raise ValueError("kin_visc must be a float/int")
        self.kin_visc = control.precision(kin_visc)
        print_log(f"{'kinematic viscosity':20s}: {self.kin_visc}", rank, verbose)
        print_log("Setting transport properties done!", rank, verbose)
        print_log("-" * 80, rank, verbose)


class Fields:
    """Per-node host arrays in the reference layout
    (pylabolt/base/fields.py:50-92): one ghost ring, ind = x * Ny_pad + y."""

    def __init__(self, control, lattice, domain, fluid=False, phase=False,
                 scalar=False):
        # same signature as the reference (base/fields.py:5-13).  The geometry
        # arrays always exist; rho / u / p are what the fluidLB path reads and
        # writes on the host (the populations live on the device only), the
        # phase-field and scalar solvers are outside this build.
        _fluid_only(phase, scalar)
        size = int(domain.size)
        prec = control.precision
        self.fluid = fluid
        self.phase = phase
        self.scalar = scalar
        self.solid = np.zeros(size, dtype=np.bool_)
        self.solid_id = np.empty(size, dtype=int)
        parallel_fill(self.solid_id.reshape(int(domain.shape[0]), -1), -1)
        self.solid_boundary = np.zeros(size, dtype=np.bool_)
        self.fluid_boundary = np.zeros(size, dtype=np.bool_)
        self.surface_normals = np.zeros((size, 2), dtype=prec)
        self.ghost_node = ghost_ring(domain.shape)
        self.periodic_boundary = np.zeros(size, dtype=np.bool_)
        self.velocity = np.zeros((size, 2), dtype=prec)
        self.density = np.zeros(size, dtype=prec)
        self.pressure = np.zeros(size, dtype=prec)


def _fluid_only(phase, scalar):
    if phase or scalar:
        raise ValueError("only the fluid (fluidLB) path is built for the b200 "
                         "back end: phase / scalar fields are not available")


def ghost_ring(shape):
    """Fields.init_ghost_nodes (pylabolt/base/fields.py:166-179), vectorised:
    True on the outermost ring of the padded array."""
    nxp, nyp = int(shape[0]), int(shape[1])
    ghost = np.zeros((nxp, nyp), dtype=np.bool_)
    ghost[0, :] = ghost[-1, :] = True
    ghost[:, 0] = ghost[:, -1] = True
    return ghost.reshape(-1)


def local_to_global(i, j, offset):
    """Local (un-padded) sub-domain index -> global index,
    parallel/cpu/MPI_kernels.py:5-18; ints or index arrays."""
    return i + offset[0], j + offset[1]


def global_to_local(i_global, j_global, offset):
    """parallel/cpu/MPI_kernels.py:21-33."""
    return i_global - offset[0], j_global - offset[1]


def global_coordinates(domain):
    """(i_global, j_global) of every padded node, flat, as int arrays:
    local_to_global(i - 1, j - 1, offset) like set_field_scalar / _vector
    (base/init_fields.py:336-345)."""
    nxp, nyp = int(domain.shape[0]), int(domain.shape[1])
    i = np.repeat(np.arange(nxp, dtype=np.int64) - 1, nyp)
    j = np.tile(np.arange(nyp, dtype=np.int64) - 1, nxp)
    offset = (int(domain.offset[0]), int(domain.offset[1]))
    return local_to_global(i, j, offset)


def _field_spec(spec, control, scalar):
    """read_dict of pylabolt/base/init_fields.py:274-322 -> (value, func)."""
    if "type" not in spec:
        raise ValueError("type missing in field definition")
    kind = spec["type"]
    if kind == "fixed":
        if "value" not in spec:
            raise ValueError("value missing for fixed type field definition")
        value = spec["value"]
        if not scalar and type(value) is list and len(value) == 2:
            return np.array(value, dtype=control.precision), None
        if scalar and type(value) in (float, int):
            return control.precision(value), None
        raise ValueError("vector value must be a list (ux, uy)" +
                         " and scalar value must be a float or int")
    if kind == "func":
        if "func" not in spec:
            raise ValueError("func missing for func type field definition")
        return None, spec["func"]
    raise ValueError("Unsupported velocity initialization")


def read_dict(dict_input, field, ghost_node, domain, control, scalar_dict=True):
    """The reference's entry point of the same name and argument order
    (base/init_fields.py:274-322): reads one field definition and applies it
    to ``field`` on the non-ghost nodes."""
    _apply_field(dict_input, field, domain, SimpleNamespace(ghost_node=ghost_node),
                 control, scalar_dict)


def set_field_scalar(domain, field, ghost_node, value=None, func=None,
                     input_file=None):
    """base/init_fields.py:352-375."""
    _set_field(field, domain, ghost_node, value, func, True, field.dtype.type)


def set_field_vector(domain, field, ghost_node, value=None, func=None,
                     input_file=None):
    """base/init_fields.py:325-349."""
    _set_field(field, domain, ghost_node, value, func, False, field.dtype.type)


def _apply_field(spec, field, domain, fields, control, scalar):
    """set_field_scalar / set_field_vector, init_fields.py:325-375.  ``func``
    is called with python ints exactly like the reference does (so the values
    are bit-identical); a function carrying ``vectorized = True`` is called
    once with index arrays instead."""
    value, func = _field_spec(spec, control, scalar)
    _set_field(field, domain, fields.ghost_node, value, func, scalar,
               control.precision)


def parallel_fill(target, value):
    """target[...] = value in row blocks on a few threads: filling a fresh
    multi-GB array is page-fault bound on one core."""
    n = target.shape[0]
    row_bytes = max(1, target[:1].nbytes)
    rows = max(1, (32 << 20) // row_bytes)
    workers = min(8, os.cpu_count() or 1, -(-n // rows))
    if workers <= 1:
        target[...] = value
        return

    def block(x0):
        target[x0:x0 + rows] = value

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(workers) as pool:
        list(pool.map(block, range(0, n, rows)))


def _set_field(field, domain, ghost_node, value, func, scalar, precision):
    """Writes `value` / `func(i_global, j_global)` to the non-ghost nodes.
    The ghost ring is by construction the outermost ring of the padded array
    (Fields.init_ghost_nodes), so the non-ghost nodes are the inner block of
    the 2-D view -- no per-node mask, no index arrays of the whole lattice."""
    nxp, nyp = int(domain.shape[0]), int(domain.shape[1])
    nx, ny = nxp - 2, nyp - 2
    ring = ghost_node.reshape(nxp, nyp)
    if ring[1:-1, 1:-1].any() or not (ring[0].all() and ring[-1].all() and
                                      ring[:, 0].all() and ring[:, -1].all()):
        return _set_field_masked(field, domain, ghost_node, value, func, scalar,
                                 precision)
    view = field.reshape((nxp, nyp) if scalar else (nxp, nyp, 2))
    inner = view[1:-1, 1:-1]
    if func is None:
        parallel_fill(inner, value)
        return
    off_i, off_j = int(domain.offset[0]), int(domain.offset[1])
    if getattr(func, "vectorized", False):
        # one call per block of rows, with flat index arrays (i, j) of equal
        # length; blocks run on a few threads (numpy releases the GIL)
        rows = max(1, min(nx, (1 << 21) // max(1, ny)))
        j_row = np.arange(ny, dtype=np.int64) + off_j

        def block(x0):
            n = min(rows, nx - x0)
            i_glob = np.repeat(np.arange(x0, x0 + n, dtype=np.int64) + off_i, ny)
            j_glob = np.tile(j_row, n)
            result = func(i_glob, j_glob)
            if scalar:
                inner[x0:x0 + n] = np.asarray(result, dtype=precision).reshape(n, ny)
            else:
                inner[x0:x0 + n, :, 0] = np.asarray(result[0], dtype=precision).reshape(n, ny)
                inner[x0:x0 + n, :, 1] = np.asarray(result[1], dtype=precision).reshape(n, ny)

        starts = range(0, nx, rows)
        workers = min(8, os.cpu_count() or 1, len(starts))
        if workers > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(workers) as pool:
                list(pool.map(block, starts))
        else:
            for x0 in starts:
                block(x0)
        return
    i_glob, j_glob = local_to_global(
        np.repeat(np.arange(nx, dtype=np.int64), ny),
        np.tile(np.arange(ny, dtype=np.int64), nx), (off_i, off_j))
    ufunc = np.frompyfunc(func, 2, 1 if scalar else 2)
    result = ufunc(i_glob.astype(object), j_glob.astype(object))
    if scalar:
        inner[...] = np.asarray(result, dtype=precision).reshape(nx, ny)
    else:
        inner[..., 0] = np.asarray(result[0], dtype=precision).reshape(nx, ny)
        inner[..., 1] = np.asarray(result[1], dtype=precision).reshape(nx, ny)


def _set_field_masked(field, domain, ghost_node, value, func, scalar, precision):
    """The same for a caller-supplied ghost mask that is not the outer ring
    (read_dict / set_field_* accept any mask, like the reference's)."""
    inner = ~ghost_node
    if func is None:
        field[inner] = value
        return
    i_glob, j_glob = global_coordinates(domain)
    i_glob, j_glob = i_glob[inner], j_glob[inner]
    if getattr(func, "vectorized", False):
        result = func(i_glob, j_glob)
    else:
        result = np.frompyfunc(func, 2, 1 if scalar else 2)(
            i_glob.astype(object), j_glob.astype(object))
    if scalar:
        field[inner] = np.asarray(result, dtype=precision)
    else:
        field[inner, 0] = np.asarray(result[0], dtype=precision)
        field[inner, 1] = np.asarray(result[1], dtype=precision)


# Sections of initial_fields_dict: (field attribute, scalar?) per key.  The
# fluid section is what the b200 path consumes; the phase section is
# initialised on whatever container carries ``phase_field`` (the reference's
# init_fields_phase, base/init_fields.py:224-268) although no phase-field
# solver is built here.
_SECTIONS = {
    "fluid": (("velocity", False), ("density", True), ("pressure", True)),
    "phase": (("phase_field", True),),
}


def _init_section(section, user, fields, domain, control, default, verbose):
    rank = domain.mpi_rank
    if default:
        for key, _ in _SECTIONS[section]:
            if key not in user:
                raise ValueError("'" + key + "' is missing in default")
    for key, scalar in _SECTIONS[section]:
        if key in user:
            if not default:
                print_log(section + ": setting " + key + " override", rank, verbose)
            _apply_field(user[key], getattr(fields, key), domain, fields, control,
                         scalar)
        else:
            print_log(section + ": no " + key + " override", rank, verbose)


def init_fields(simulation, control, domain, fields, fluid=False, phase=False,
                scalar=False, verbose=True):
    """initial_fields_dict -- pylabolt/base/init_fields.py:7-129 (same
    signature): the ``default`` section is mandatory, every other key is a
    region override applied in dict order."""
    rank = domain.mpi_rank
    print_log("-" * 80, rank, verbose)
    print_log("Initializing fields...\n", rank, verbose)
    initial_fields_dict = _require(simulation, "initial_fields_dict")
    if "default" not in initial_fields_dict:
        raise ValueError("default missing in initial_fields_dict")
    default = initial_fields_dict["default"]
    enabled = (("fluid", fluid, "fluid missing in initial_fields_dict - default",
                "WARNING! not a fluid solver. Skipping fluid overrides"),
               ("phase", phase, "'phase' missing in initial_fields_dict - default",
                "WARNING! not a multiphase solver. Skipping phase overrides"))
    for section, on, missing, _ in enabled:
        if on is True:
            if section not in default:
                raise ValueError(missing)
            _init_section(section, default[section], fields, domain, control,
                          True, verbose)
    for region_no, (region, user) in enumerate(initial_fields_dict.items()):
        if region == "default":
            continue
        print_log("Region id: " + str(region_no) + " | Region name: " +
                  str(region), rank, verbose)
        for section, on, _, skipped in enabled:
            if section not in user:
                print_log(section + ": no override", rank, verbose)
            elif on is False:
                print_log(skipped, rank, verbose)
            else:
                print_log(section + ": override present", rank, verbose)
                _init_section(section, user[section], fields, domain, control,
                              False, verbose)
    print_log("Initializing fields done!", rank, verbose)
    print_log("-" * 80, rank, verbose)


class State:
    """Everything the step needs, built in the reference's order
    (pylabolt/base/state.py:33-113): Control, Mesh, Lattice, Domain,
    Transport, Fields, init_fields, Boundary, Obstacle.  Configuration errors
    print ``FATAL ERROR!`` on rank 0 and abort the communicator, like the
    reference (:121-125)."""

    def __init__(self, simulation, comm, mpi_rank=0, fluid=True, verbose=True):
        from .boundary import Boundary
        from .obstacle import Obstacle
        self.fluid, self.phase, self.scalar = True, False, False
        try:
            self.control = Control(simulation, mpi_rank, verbose)
            if self.control.precision is not np.float64:
                raise ValueError(
                    "precision 'single' is not available in the b200 back end: "
                    "the fluidLB step is an fp64 path (use 'double')")
            self.mesh = Mesh(simulation, mpi_rank, verbose)
            self.lattice = Lattice(simulation, self.control, self.mesh,
                                   mpi_rank, verbose)
            if self.lattice.lattice_type != "D2Q9":
                raise ValueError(self.lattice.lattice_type + " is not available "
                                 "in the b200 back end (the accelerated path "
                                 "is D2Q9)")
            self.domain = Domain(simulation, self.mesh, comm, verbose)
            self.domain.require_slabs()
            self.transport = Transport(simulation, self.control, self.domain,
                                       verbose)
            self.fields = Fields(self.control, self.lattice, self.domain,
                                 fluid=fluid)
            init_fields(simulation, self.control, self.domain, self.fields,
                        fluid=fluid, verbose=verbose)
            self.boundary = Boundary(simulation, self.mesh, self.domain,
                                     self.control, self.fields, fluid=fluid,
                                     verbose=verbose)
            self.obstacle = Obstacle(simulation, self.mesh, self.domain,
                                     self.control, self.fields, self.boundary,
                                     fluid=fluid, verbose=verbose)
        except Exception as e:
            print_log("-" * 80, mpi_rank, True)
            print_log("FATAL ERROR!", mpi_rank, True)
            print_log(str(e), mpi_rank, True)
            comm.Abort()
            raise
