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
import numpy as np

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
            raise ValueError(
                "precision 'single' is not available in the b200 back end: "
                "the fluidLB step is an fp64 path (use 'double')")
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
        if self.lattice_type == "D2Q9":
            if mesh.dimensions != 2:
                raise ValueError(
                    "grid dimensions and lattice type are incompatible")
        elif self.lattice_type == "D1Q3":
            raise ValueError("D1Q3 is not available in the b200 back end "
                             "(the accelerated path is D2Q9)")
        else:
            raise ValueError("Unsupported lattice type")
        for key, value in d2q9_constants(control.precision).items():
            setattr(self, key, value)
        self.no_of_directions = int(9)
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

    def __init__(self, control, lattice, domain):
        size = int(domain.size)
        prec = control.precision
        self.fluid = True
        self.solid = np.zeros(size, dtype=np.bool_)
        self.solid_id = np.full(size, -1, dtype=int)
        self.solid_boundary = np.zeros(size, dtype=np.bool_)
        self.fluid_boundary = np.zeros(size, dtype=np.bool_)
        self.surface_normals = np.zeros((size, 2), dtype=prec)
        self.ghost_node = ghost_ring(domain.shape)
        self.periodic_boundary = np.zeros(size, dtype=np.bool_)
        self.velocity = np.zeros((size, 2), dtype=prec)
        self.density = np.zeros(size, dtype=prec)
        self.pressure = np.zeros(size, dtype=prec)


def ghost_ring(shape):
    """Fields.init_ghost_nodes (pylabolt/base/fields.py:166-179), vectorised:
    True on the outermost ring of the padded array."""
    nxp, nyp = int(shape[0]), int(shape[1])
    ghost = np.zeros((nxp, nyp), dtype=np.bool_)
    ghost[0, :] = ghost[-1, :] = True
    ghost[:, 0] = ghost[:, -1] = True
    return ghost.reshape(-1)


def global_coordinates(domain):
    """(i_global, j_global) of every padded node, flat, as int arrays:
    local_to_global(i - 1, j - 1, offset), cpu/MPI_kernels.py:5-18."""
    nxp, nyp = int(domain.shape[0]), int(domain.shape[1])
    i = np.repeat(np.arange(nxp, dtype=np.int64) - 1 + int(domain.offset[0]), nyp)
    j = np.tile(np.arange(nyp, dtype=np.int64) - 1 + int(domain.offset[1]), nxp)
    return i, j


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


def _apply_field(spec, field, domain, fields, control, scalar):
    """set_field_scalar / set_field_vector, init_fields.py:325-375.  ``func``
    is called with python ints exactly like the reference does (so the values
    are bit-identical); a function carrying ``vectorized = True`` is called
    once with index arrays instead."""
    value, func = _field_spec(spec, control, scalar)
    inner = ~fields.ghost_node
    if func is None:
        field[inner] = value
        return
    i_glob, j_glob = global_coordinates(domain)
    i_glob, j_glob = i_glob[inner], j_glob[inner]
    if getattr(func, "vectorized", False):
        result = func(i_glob, j_glob)
    elif scalar:
        result = np.frompyfunc(func, 2, 1)(i_glob.astype(object),
                                           j_glob.astype(object))
    else:
        result = np.frompyfunc(func, 2, 2)(i_glob.astype(object),
                                           j_glob.astype(object))
    if scalar:
        field[inner] = np.asarray(result, dtype=control.precision)
    else:
        field[inner, 0] = np.asarray(result[0], dtype=control.precision)
        field[inner, 1] = np.asarray(result[1], dtype=control.precision)


def init_fields(simulation, control, domain, fields, verbose=True):
    """initial_fields_dict -- pylabolt/base/init_fields.py:7-129: the
    ``default`` section is mandatory, every other key is a region override
    applied in dict order."""
    rank = domain.mpi_rank
    print_log("-" * 80, rank, verbose)
    print_log("Initializing fields...\n", rank, verbose)
    initial_fields_dict = _require(simulation, "initial_fields_dict")
    if "default" not in initial_fields_dict:
        raise ValueError("default missing in initial_fields_dict")
    default = initial_fields_dict["default"]
    if "fluid" not in default:
        raise ValueError("fluid missing in initial_fields_dict - default")
    fluid = default["fluid"]
    for key in ("velocity", "density", "pressure"):
        if key not in fluid:
            raise ValueError("'" + key + "' is missing in default")
    targets = (("velocity", fields.velocity, False),
               ("density", fields.density, True),
               ("pressure", fields.pressure, True))
    for key, field, scalar in targets:
        _apply_field(fluid[key], field, domain, fields, control, scalar)
    for region_no, (region, user) in enumerate(initial_fields_dict.items()):
        if region == "default":
            continue
        print_log("Region id: " + str(region_no) + " | Region name: " +
                  str(region), rank, verbose)
        if "fluid" not in user:
            print_log("fluid: no override", rank, verbose)
            continue
        print_log("fluid: override present", rank, verbose)
        for key, field, scalar in targets:
            if key in user["fluid"]:
                print_log("fluid: setting " + key + " override", rank, verbose)
                _apply_field(user["fluid"][key], field, domain, fields,
                             control, scalar)
            else:
                print_log("fluid: no " + key + " override", rank, verbose)
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
            self.mesh = Mesh(simulation, mpi_rank, verbose)
            self.lattice = Lattice(simulation, self.control, self.mesh,
                                   mpi_rank, verbose)
            self.domain = Domain(simulation, self.mesh, comm, verbose)
            self.domain.require_slabs()
            self.transport = Transport(simulation, self.control, self.domain,
                                       verbose)
            self.fields = Fields(self.control, self.lattice, self.domain)
            init_fields(simulation, self.control, self.domain, self.fields,
                        verbose)
            self.boundary = Boundary(simulation, self.mesh, self.domain,
                                     self.control, self.fields, verbose)
            self.obstacle = Obstacle(simulation, self.mesh, self.domain,
                                     self.control, self.fields, self.boundary,
                                     verbose)
        except Exception as e:
            print_log("-" * 80, mpi_rank, True)
            print_log("FATAL ERROR!", mpi_rank, True)
            print_log(str(e), mpi_rank, True)
            comm.Abort()
            raise
