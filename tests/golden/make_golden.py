#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the REFERENCE
itself (imported read-only from /root/reference) on the cases of
tests/cases.py.

Run from the repo root, in the build container only (the GPU box has no
/root/reference):

    python tests/golden/make_golden.py

What runs: the reference's own containers (State: Control, Mesh, Lattice,
Domain, Transport, Fields, init_fields, Boundary, Obstacle), its own operators
and its own numba CPU kernels, in the order of
pylabolt/solvers/fluidLB.py:206-253 (Solver.single_time_step).

Shims (none changes arithmetic):
  1. mpi4py is not installed -> a single-rank stand-in module.
  2. importlib.metadata.version("pylabolt") -> "1.0.0.dev0" (not installed).
  3. ObstacleOperator.{set_backend, compile, move_obstacles,
     compute_force_torque} are no-ops: upstream crashes at
     base/obstacle_operator.py:516 and :67; a no-op is exact for bodies whose
     mask does not change.
  4. BoundaryElement.surface_normals is cast to int64 before the
     fixed_pressure kernel is compiled (float index typing error,
     cpu/fluid_boundary_kernels.py:133-137).
  5. Single-rank periodic ghosts are filled with a true WRAP
     (gpu/MPI_kernels.py:34-99 semantics; multi-rank CPU semantics) instead of
     the self-Sendrecv mirror of MPI_operator.py:261-314.

Each fixture <case>.npz stores: flags and ids after setup, every boundary
element's link list, initial rho/u, and rho / u / pop_new (padded reference
layouts) after the recorded step counts, the wall / body forces of the
reference's force kernels, and the residue sums + residues of the reference's
residue operator called at exactly the recorded steps (field_old = 0 before
the first).  Fixtures are small (a few 100 KB).
"""
import contextlib
import importlib.metadata
import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"


# --------------------------------------------------------------------------
# shims 1 + 2
# --------------------------------------------------------------------------
class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Barrier(self):
        pass

    def Abort(self, code=1):
        raise RuntimeError("comm.Abort() called by the reference")

    def Sendrecv(self, sendbuf, dest, sendtag, recvbuf, source, recvtag):
        recvbuf[...] = sendbuf

    def Allreduce(self, local, glob, op=None):
        glob[...] = local


def install_shims():
    mpi4py = types.ModuleType("mpi4py")
    mpi4py.rc = lambda **kw: None
    mpi = types.ModuleType("mpi4py.MPI")
    mpi.COMM_WORLD = _Comm()
    mpi.SUM = "sum"
    mpi.Init = lambda: None
    mpi.Finalize = lambda: None
    mpi4py.MPI = mpi
    sys.modules["mpi4py"] = mpi4py
    sys.modules["mpi4py.MPI"] = mpi
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    real_version = importlib.metadata.version

    def version(name):
        if name == "pylabolt":
            return "1.0.0.dev0"
        return real_version(name)
    importlib.metadata.version = version
    return mpi.COMM_WORLD


def wrap_ghosts(field, shape, x_periodic, y_periodic):
    """shim 5: x-phase (full columns) then y-phase (full rows)."""
    nxp, nyp = int(shape[0]), int(shape[1])
    view = field.reshape((nxp, nyp) + field.shape[1:])
    if x_periodic:
        view[0] = view[nxp - 2]
        view[nxp - 1] = view[1]
    if y_periodic:
        view[:, 0] = view[:, nyp - 2]
        view[:, nyp - 1] = view[:, 1]


def build_reference_solver(simulation, comm, n_threads=4):
    """Reference Solver with shims 3-5 applied; returns it compiled and with
    the populations initialised (run() up to the start of the time loop,
    solvers/fluidLB.py:309-343, minus I/O)."""
    from pylabolt.solvers import fluidLB
    from pylabolt.utils import helpers
    from pylabolt.base.obstacle_operator import ObstacleOperator
    from pylabolt.parallel.MPI_operator import MPIOperator

    helpers.load_simulation = lambda comm, rank: simulation
    fluidLB.load_simulation = helpers.load_simulation

    # shim 5
    def halo_exchange_wrap(self, state, backend, bool_buffers=None,
                           int_buffers=None, float_buffers=None):
        for names in (bool_buffers, int_buffers, float_buffers):
            if names is None:
                continue
            for name in names:
                wrap_ghosts(getattr(state.fields, name), state.domain.shape,
                            state.boundary.x_periodic,
                            state.boundary.y_periodic)
    MPIOperator.halo_exchange_cpu = halo_exchange_wrap

    # shim 3
    ObstacleOperator.set_backend = lambda self, state, backend: None
    ObstacleOperator.compile = lambda self, state, backend: None
    ObstacleOperator.move_obstacles = lambda self, *a, **k: None
    ObstacleOperator.compute_force_torque = lambda self, *a, **k: None
    ObstacleOperator.verify_kernel_signatures = lambda self, *a, **k: None

    solver = fluidLB.Solver(comm, "cpu", n_threads)
    # shim 4
    for element in solver.state.boundary.boundary_elements:
        element.surface_normals = element.surface_normals.astype(np.int64)
    solver.mpi_operator.halo_exchange = types.MethodType(
        halo_exchange_wrap, solver.mpi_operator)
    solver.set_backend(verbose=False)
    solver.mpi_operator.halo_exchange = types.MethodType(
        halo_exchange_wrap, solver.mpi_operator)
    solver.compile(verbose=False)
    solver.collision_operator.initialize_pop(solver.state, solver.backend)
    return solver


def snapshot_setup(solver):
    st = solver.state
    f = st.fields
    out = {
        "shape": np.asarray(st.domain.shape, dtype=np.int64),
        "solid": f.solid.copy(),
        "ghost_node": f.ghost_node.copy(),
        "solid_id": f.solid_id.copy(),
        "solid_boundary": f.solid_boundary.copy(),
        "fluid_boundary": f.fluid_boundary.copy(),
        "periodic_boundary": f.periodic_boundary.copy(),
        "surface_normals": f.surface_normals.copy(),
        "density_0": f.density.copy(),
        "velocity_0": f.velocity.copy(),
        "pop_0": f.pop_fluid_new.copy(),
        "omega": np.float64(solver.collision_operator.omega_fluid),
        "gravity": np.asarray(solver.force_operator.gravity, np.float64),
        "x_periodic": np.bool_(st.boundary.x_periodic),
        "y_periodic": np.bool_(st.boundary.y_periodic),
        "n_elements": np.int64(len(st.boundary.boundary_elements)),
        "lattice_consts": np.array([st.lattice.cs, st.lattice.cs_2,
                                    st.lattice.inv_cs_2, st.lattice.inv_cs_4,
                                    st.control.float_min], np.float64),
        "weights": st.lattice.weights.copy(),
    }
    for n, el in enumerate(st.boundary.boundary_elements):
        out[f"el{n}_name"] = np.array(el.name)
        out[f"el{n}_type"] = np.array(el.type_fluid)
        out[f"el{n}_nodes"] = el.boundary_nodes.copy()
        out[f"el{n}_out"] = el.out_list.copy()
        out[f"el{n}_inv"] = el.inv_list.copy()
        out[f"el{n}_normal"] = np.asarray(el.surface_normals, np.int64)
        out[f"el{n}_vector"] = np.asarray(el.vector_fluid, np.float64)
        out[f"el{n}_scalar"] = np.float64(el.scalar_fluid)
    return out


def reference_forces(solver):
    """Wall and obstacle forces computed by the reference's own kernels
    (cpu/force_torque_kernels.py) on the current pop / pop_new."""
    from pylabolt.parallel.cpu.force_torque_kernels import (
        compute_boundary_force_single_phase,
        compute_force_torque_single_phase)
    st = solver.state
    f, lat = st.fields, st.lattice
    wall = np.array([
        compute_boundary_force_single_phase(
            lat.cx, lat.cy, f.solid, f.pop_fluid, f.pop_fluid_new,
            el.boundary_nodes, el.out_list, el.inv_list)
        for el in st.boundary.boundary_elements]).reshape(-1, 2)
    body = np.array([
        compute_force_torque_single_phase(
            st.domain.size, st.domain.shape, st.domain.offset,
            st.mesh.grid_global_shape, lat.cx, lat.cy, lat.inv_list,
            lat.no_of_directions, st.boundary.x_periodic,
            st.boundary.y_periodic, f.solid, f.solid_id, f.fluid_boundary,
            f.ghost_node, f.pop_fluid, f.pop_fluid_new, ob.ref_point, ob.id)
        for ob in st.obstacle.obstacles]).reshape(-1, 3)
    return wall, body


def reference_residues(solver, time_step):
    """The reference's own residue operator (utils/residues.py:171-222 with
    the numba kernels of cpu/compute_residues_kernels.py:6-73) at a recorded
    step: returns the raw sums it reduces -- (num, den) of density and
    (num_x, den_x, num_y, den_y) of velocity -- and the residues it logs.
    field_old of the operator lives on between calls (zeros before the first
    one), exactly as in Solver.run with std_out_interval = the recorded
    steps."""
    st = solver.state
    op = solver.residue_operator
    seen = []
    real_reduce = solver.mpi_operator.reduce

    def reduce(local_array, operation="sum"):
        seen.append(np.array(local_array, dtype=np.float64))
        return real_reduce(local_array, operation=operation)
    keep = st.control.std_out_interval
    st.control.std_out_interval = 1
    solver.mpi_operator.reduce = reduce
    try:
        op.compute_residues(st, solver.backend, solver.mpi_operator, time_step)
    finally:
        solver.mpi_operator.reduce = real_reduce
        st.control.std_out_interval = keep
    order = list(op.fields_list)
    sums = np.concatenate([seen[order.index("density")],
                           seen[order.index("velocity")]])
    res = np.concatenate([op.residues["res_density"],
                          op.residues["res_velocity"]])
    return sums, np.asarray(res, np.float64)


def generate(case_name, factory, kwargs, record_steps, comm, out_dir):
    simulation = factory(**kwargs)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)      # the reference writes metadata.json / output/ in cwd
        try:
            with contextlib.redirect_stdout(io.StringIO()):   # setup chatter
                solver = build_reference_solver(simulation, comm)
            data = snapshot_setup(solver)
            fields = solver.state.fields
            if case_name == "cylinder":
                # what the reference writes at t = start_time
                # (utils/io_operator.py:158-190 + dump_metadata :96-156)
                solver.state.control.save_interval = 1
                solver.io_operator.write_fields(solver.state, solver.backend, 0)
                solver.state.control.save_interval = None
                saved = np.load(os.path.join("output", "fields", "t_0.npz"))
                for key in saved.files:
                    data["io_" + key] = saved[key]
                data["io_metadata_json"] = np.array(
                    open("metadata.json").read())
            data["wall_force_0"], data["body_force_0"] = \
                reference_forces(solver)
            for step in range(1, max(record_steps) + 1):
                solver.single_time_step()
                if step in record_steps:
                    data[f"wall_force_{step}"], data[f"body_force_{step}"] = \
                        reference_forces(solver)
                    data[f"density_{step}"] = fields.density.copy()
                    data[f"velocity_{step}"] = fields.velocity.copy()
                    data[f"pop_{step}"] = fields.pop_fluid_new.copy()
                    data[f"residue_sums_{step}"], data[f"residues_{step}"] = \
                        reference_residues(solver, step)
            data["record_steps"] = np.asarray(record_steps, np.int64)
        finally:
            os.chdir(cwd)
    path = os.path.join(out_dir, case_name + ".npz")
    np.savez_compressed(path, **data)
    mass = float(data[f"density_{max(record_steps)}"]
                 [~data["ghost_node"] & ~data["solid"]].sum())
    print(f"[golden] {case_name:<18} shape={tuple(int(v) for v in data['shape'])} "
          f"steps={tuple(record_steps)} fluid mass={mass:.10f} -> "
          f"{os.path.relpath(path, REPO)} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import cases
    comm = install_shims()
    only = set(sys.argv[1:])
    for name, (factory, kwargs, steps) in cases.GOLDEN_CASES.items():
        if only and name not in only:
            continue
        generate(name, factory, kwargs, steps, comm, HERE)


if __name__ == "__main__":
    main()
