"""The fluidLB solver on the b200 back end: mirror of
pylabolt/solvers/fluidLB.py (Solver :108-395, main :398-407).

Same life cycle -- ``Solver(comm, backend, n_threads)``, ``set_backend()``,
``compile()``, ``run()`` -- and the same seam: ``execute_single_time_step`` is
a zero-argument callable bound in ``compile()`` (reference :273-280) and
called once per step by ``run()`` (:356).  In the reference that callable runs
nine operator calls and five full-lattice numba kernels; here it is one call
into libplb (``plb_step``), which fuses phases 2-8 into a single pass over HBM.

There is no CPU path: without the CUDA library or a CUDA device the solver
raises.
"""
import os
import sys
import time

import numpy as np

from . import capi
from .force_torque import MomentumExchange
from .helpers import SimulationStatusLogger, load_simulation, print_log
from .io_operator import InputOutputOperator
from .operators import CollisionOperator, FluidLB, ForceOperator
from .residues import ResidueOperator
from .state import State

__version__ = "1.0.0.dev0+b200"


class Backend:
    """Hardware back end descriptor (pylabolt/parallel/backend.py:18-81).
    The only back end of this package is ``b200``: one process per GPU."""

    def __init__(self, comm, state, backend="b200", n_threads=1, device=None,
                 strict=None, verbose=True):
        rank = state.domain.mpi_rank
        if backend not in ("b200", "gpu"):
            print_log("-" * 80, rank, True)
            print_log("FATAL ERROR!", rank, True)
            print_log("pylabolt_b200 has no '" + str(backend) + "' back end; "
                      "use --backend b200", rank, True)
            comm.Abort()
            raise ValueError("unsupported backend: " + str(backend))
        self.backend_type = "b200"
        self.no_of_threads = n_threads
        self.strict = strict
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        capi.load_library(strict)      # fail loudly now, not at the first step
        # run (and allocate transfer buffers) on the GPU's own CPU socket
        self.numa_node = None
        if state.domain.mpi_size > 1:
            from .affinity import bind_to_gpu
            try:
                self.numa_node = bind_to_gpu(
                    capi.device_pci_bus_id(self.device, strict))
            except capi.PlbError:
                pass                   # plb_create reports the missing device
            if os.environ.get("PLB_DEBUG"):
                print(f"[plb] rank {rank} device {self.device}: NUMA node "
                      f"{self.numa_node}, cpus "
                      f"{len(os.sched_getaffinity(0))}", file=sys.stderr,
                      flush=True)
        print_log(f"{'Backend':<25}: b200 (sm_100a, libplb)", rank, verbose)
        print_log(f"{'CUDA device':<25}: {self.device}", rank, verbose)


def neighbour_ranks(domain, boundary):
    """MPIOperator.find_neighbor_ranks (parallel/MPI_operator.py:116-153) for
    x-slabs: (left_rank, right_rank), None where the slab ends at a
    non-periodic domain edge."""
    n = domain.no_of_procs_x
    i = domain.i_proc
    left = None if (i == 0 and not boundary.x_periodic) else (i - 1 + n) % n
    right = None if (i == n - 1 and not boundary.x_periodic) else (i + 1) % n
    return left, right


class Solver:
    def __init__(self, comm, backend="b200", n_threads=1, simulation=None,
                 device=None, strict=None, verbose=True):
        mpi_rank = comm.Get_rank()
        self.comm = comm
        self.verbose = verbose
        print_log(f"\n{'PyLaBolt':<10}: {__version__}", mpi_rank, verbose)
        print_log(f"{'Solver':<10}: fluidLB", mpi_rank, verbose)
        self.model = FluidLB()
        if simulation is None:
            simulation = load_simulation(comm, mpi_rank)
        self.simulation = simulation
        self.state = State(simulation, comm, mpi_rank, fluid=True,
                           verbose=verbose)
        self.backend = Backend(comm, self.state, backend, n_threads,
                               device=device, strict=strict, verbose=verbose)
        self.collision_operator = CollisionOperator(
            simulation, self.model, self.state, comm, verbose=verbose)
        self.force_operator = ForceOperator(
            simulation, self.model, self.state, comm,
            collision_operator=self.collision_operator, verbose=verbose)
        try:
            self.state.obstacle.check_overlap(comm)
        except RuntimeError as e:
            print_log(str(e), mpi_rank, True)
            comm.Abort()
            raise
        self.residue_operator = ResidueOperator(self.model, self.state, comm,
                                                verbose=verbose)
        self.io_operator = InputOutputOperator(self.model, self.state,
                                               self.backend, comm,
                                               verbose=verbose)
        self.logger = SimulationStatusLogger(mpi_rank, verbose=verbose)
        self.plb = None
        self.time_step = self.state.control.start_time
        self.execute_single_time_step = None

    # -- reference: Solver.set_backend, fluidLB.py:186-204 ---------------------
    def set_backend(self, verbose=None):
        verbose = self.verbose if verbose is None else verbose
        st = self.state
        rank = st.domain.mpi_rank
        print_log("-" * 80, rank, verbose)
        print_log("Setting simulation backend...\n", rank, verbose)
        left, right = neighbour_ranks(st.domain, st.boundary)
        self.left_rank, self.right_rank = left, right
        col = self.collision_operator
        self.plb = capi.Plb(
            st.domain.Nx_rank, st.domain.Ny_rank, col.omega_fluid,
            device=self.backend.device, collision=col.collision_fluid,
            forcing=col.forcing_fluid, gravity=self.force_operator.gravity,
            x_periodic=st.boundary.x_periodic,
            y_periodic=st.boundary.y_periodic,
            left_neighbor=left is not None, right_neighbor=right is not None,
            mrt_rates=col.mrt_rates,
            lattice={"inv_cs_2": st.lattice.inv_cs_2,
                     "inv_cs_4": st.lattice.inv_cs_4,
                     "weights": st.lattice.weights},
            float_min=st.control.float_min, strict=self.backend.strict)
        plb = self.plb
        plb.upload(capi.SOLID, st.fields.solid)
        self.upload_initial_fields()
        for element in st.boundary.boundary_elements:
            plb.add_boundary_element(
                element.type_fluid, element.boundary_nodes, element.out_list,
                element.inv_list, element.normal, element.vector_fluid,
                element.scalar_fluid)
        plb.finalize_geometry()
        if st.domain.mpi_size > 1:
            unique_id = plb.comm_unique_id() if rank == 0 else bytes(128)
            unique_id = self.comm.bcast_bytes(unique_id, root=0)
            plb.comm_init(unique_id, rank, st.domain.mpi_size,
                          -1 if left is None else left,
                          -1 if right is None else right)
        self.residue_operator.set_backend(st, self.backend, plb)
        self.io_operator.set_backend(st, self.backend, plb)
        self.momentum = None
        if st.boundary.compute_force or st.obstacle.compute_force_torque:
            self.momentum = MomentumExchange(st, plb.link_nodes())
        print_log("\nSetting simulation backend done!", rank, verbose)
        print_log("-" * 80, rank, verbose)

    def _uniform_initial_value(self, key):
        """The value of initial field ``key`` ("density" / "velocity") if the
        case file fixes it to ONE value on every fluid node -- ``type: fixed``
        in the default section, no region override -- and no obstacle writes
        its own density / velocity into solid nodes; else None."""
        if self.state.obstacle.obstacles:
            return None
        regions = getattr(self.simulation, "initial_fields_dict", {})
        spec = regions.get("default", {}).get("fluid", {}).get(key)
        if not isinstance(spec, dict) or spec.get("type") != "fixed":
            return None
        for name, user in regions.items():
            if name != "default" and key in user.get("fluid", {}):
                return None
        return spec["value"]

    def upload_initial_fields(self, density=None, velocity=None):
        """fields.density / fields.velocity -> device (State.set_backend of
        the reference mirrors every field array, base/fields.py:192-227).  A
        field the case file fixes to one value everywhere is produced on the
        device (plb_fill) instead of crossing PCIe as 8 (16) bytes per node;
        everything else is uploaded from ``density`` / ``velocity`` (default:
        the State's own arrays; a caller may pass pinned copies).  Returns
        the bytes that crossed PCIe."""
        st = self.state
        moved = 0
        for key, field, host in (("density", capi.DENSITY, density),
                                 ("velocity", capi.VELOCITY, velocity)):
            value = self._uniform_initial_value(key)
            if value is not None:
                self.plb.fill(field, value)
                continue
            if host is None:
                host = getattr(st.fields, key)
            self.plb.upload(field, host)
            moved += host.nbytes
        return moved

    # -- reference: Solver.single_time_step, fluidLB.py:206-253 ----------------
    def single_time_step(self, store_moments=False, record_links=False):
        """One reference time step: phases 2-8 (moments, gravity force,
        collision, halo exchange, streaming with in-flight bounce back,
        boundary elements) as one fused pass inside libplb.  With
        ``store_moments`` the step also leaves rho / u (phases 2-4) in HBM for
        output and residues, which the reference does on every step; with
        ``record_links`` it records the momentum exchanged on wall / solid
        links (phase 9 and BoundaryOperator.compute_force)."""
        self.plb.step(1, store_moments, record_links)

    def advance(self, n_steps, store_moments_last=False, record_links_last=False):
        """``n_steps`` calls of single_time_step in one library call."""
        self.plb.step(n_steps, store_moments_last, record_links_last)

    # -- reference: BoundaryOperator.compute_force (boundary_operator.py:204-237)
    #    and ObstacleOperator.compute_force_torque (obstacle_operator.py:75-109)
    def compute_forces(self, exchange=None):
        """Reduces the recorded link exchange to the force on every wall
        boundary element and the force / torque on every obstacle (summed
        over ranks) and stores them on the element / obstacle objects."""
        st = self.state
        if self.momentum is None:
            return None, None
        if exchange is None:
            exchange = self.plb.link_exchange(self.momentum.n_links)
        wall_local = self.momentum.boundary_forces(exchange)
        body_local = self.momentum.obstacle_forces(exchange)
        wall = np.zeros_like(wall_local)
        body = np.zeros_like(body_local)
        self.comm.Allreduce(wall_local, wall)
        self.comm.Allreduce(body_local, body)
        for n, element in enumerate(st.boundary.boundary_elements):
            if element.wall:
                element.force[:] = wall[n]
        st.boundary.local_force[:] = wall_local
        st.boundary.global_force[:] = wall
        for n, body_obj in enumerate(st.obstacle.obstacles):
            body_obj.force[:] = body[n, :2]
            body_obj.torque = body[n, 2]
        return wall, body

    def _needs_links(self, time_step):
        st = self.state
        if self.momentum is None:
            return False
        wall = (st.boundary.compute_force and
                st.boundary.write_interval is not None and
                time_step % st.boundary.write_interval == 0)
        body = (st.obstacle.compute_force_torque and
                st.obstacle.write_interval is not None and
                time_step % st.obstacle.write_interval == 0)
        return wall or body

    # -- reference: Solver.compile, fluidLB.py:255-284 -------------------------
    def compile(self, verbose=None):
        """Nothing is JIT-compiled: libplb is built ahead of time for sm_100a.
        Binds the step slot, like the reference binds it to its CUDA graph."""
        verbose = self.verbose if verbose is None else verbose
        if self.plb is None:
            raise RuntimeError("set_backend() must be called before compile()")
        self.execute_single_time_step = self.single_time_step
        print_log("Step slot bound to libplb (plb_step)",
                  self.state.domain.mpi_rank, verbose)

    def _needs_moments(self, time_step):
        c = self.state.control
        return ((c.std_out_interval is not None and
                 time_step % c.std_out_interval == 0) or
                (c.save_interval is not None and
                 time_step % c.save_interval == 0))

    # -- reference: Solver.run, fluidLB.py:309-395 ------------------------------
    def run(self, verbose=None):
        verbose = self.verbose if verbose is None else verbose
        st = self.state
        rank = st.domain.mpi_rank
        print_log("\n" + "-" * 80, rank, verbose)
        print_log("Running simulation...\n", rank, verbose)
        # Restart (start_time > 0 with a checkpoint of that step): a decision
        # of the whole job -- a rank whose file is missing must not quietly
        # begin from f_eq while its neighbours resume.
        found = self.io_operator.has_checkpoint(st, st.control.start_time)
        resumed = self.comm.all_agree(found)
        if found and not resumed:
            print_log("-" * 80, rank, True)
            print_log("FATAL ERROR!", rank, True)
            print_log("checkpoint of step " + str(st.control.start_time) +
                      " is missing on at least one rank", rank, True)
            self.comm.Abort()
            raise RuntimeError("checkpoint missing on a neighbouring rank")
        if resumed:
            # The populations of step start_time come back; that step's
            # fields, histories and forces were written by the run that made
            # the checkpoint (rho / u on the device still hold the case
            # file's initial fields until the next output step), so nothing
            # is written for it again and the history files are continued.
            self.io_operator.load_checkpoint(st, st.control.start_time)
            self.io_operator.resume_histories(st)
        else:
            self.plb.initialize_pop()
            if self.momentum is not None:
                self.compute_forces(self.momentum.initial_exchange(st.lattice))
            self.io_operator.write_fields(st, self.backend,
                                          st.control.start_time)
            self.io_operator.write_histories(st, time_step=0)
        run_time_start = time.perf_counter()
        for time_step in range(st.control.start_time + 1,
                               st.control.end_time + 1):
            moments = self._needs_moments(time_step)
            links = self._needs_links(time_step)
            if moments or links:
                self.single_time_step(store_moments=moments, record_links=links)
            else:
                self.execute_single_time_step()
            self.time_step = time_step
            if links:
                self.compute_forces()
            self.residue_operator.compute_residues(st, self.backend, self.comm,
                                                   time_step)
            self.logger.log_data(
                st, time_step,
                res_density=self.residue_operator.residues["res_density"],
                res_velocity=self.residue_operator.residues["res_velocity"])
            self.io_operator.write_fields(st, self.backend, time_step)
            self.io_operator.write_histories(st, time_step)
            self.io_operator.write_checkpoint(st, time_step)
            if moments or links:
                # data left the device on this step: a slab-face time-out
                # (dead neighbour rank) must stop the run here, not after
                # the last step
                self.plb.sync()
        self.plb.sync()
        run_time = time.perf_counter() - run_time_start
        print_log("\n" + "-" * 80, rank, verbose)
        print_log(f"{'Simulation complete, run time':<30}: "
                  f"{str(run_time) + ' s':<30}", rank, verbose)
        print_log("-" * 80, rank, verbose)
        return run_time

    # -- host views of device fields (parity tests, post-processing) ------------
    def fields_to_host(self):
        """density, velocity and pop_fluid_new in the reference's padded
        layouts, as numpy arrays."""
        return {"density": self.plb.download(capi.DENSITY),
                "velocity": self.plb.download(capi.VELOCITY),
                "pop_fluid_new": self.plb.download(capi.POP)}

    def close(self):
        self.io_operator.close()
        if self.plb is not None:
            self.plb.close()
            self.plb = None


def main(backend="b200", n_threads=1, debug_mode=False):
    """pylabolt/solvers/fluidLB.py:398-407."""
    from .comm import world_comm
    comm = world_comm()
    solver = Solver(comm, backend, n_threads)
    solver.set_backend()
    solver.compile()
    solver.run()
    solver.close()
