"""Field output: mirror of pylabolt/utils/io_operator.py.

Writes what the reference writes, where the reference writes it, so that its
``--reconstruct`` and ``--to_vtk`` post-processing keep working unchanged:

* ``metadata.json`` (io_operator.py:96-156),
* ``output/fields/t_<n>.npz`` on one rank, ``procs/proc_<r>/t_<n>.npz`` plus
  ``rank_metadata.json`` on several (io_operator.py:82-95, 158-190), each
  holding the arrays of FluidLB.save_fields with the ghost ring stripped,
  x-major.

rho and u come from the device (plb_download of the inner region); the flag
fields never change for static bodies and are sliced from the host arrays.
"""
import json
import os

import numpy as np

from . import capi
from .helpers import print_log


def strip_ghost(field, shape):
    """copy_inner_data_{scalar,vector}, cpu/io_operator_kernels.py:5-52."""
    nxp, nyp = int(shape[0]), int(shape[1])
    view = field.reshape((nxp, nyp) + field.shape[1:])[1:-1, 1:-1]
    return np.ascontiguousarray(view).reshape((-1,) + field.shape[1:])


class InputOutputOperator:
    def __init__(self, model, state, backend, comm, verbose=True,
                 root_dir="."):
        rank = state.domain.mpi_rank
        print_log("-" * 80, rank, verbose)
        print_log("Setting up I/O operator...\n", rank, verbose)
        self.model = model
        self.plb = None
        self.root_dir = root_dir
        self.fields_list = model.save_fields
        self.fields_save_metadata = {}
        for name in self.fields_list:
            if not hasattr(state.fields, name):
                raise ValueError(name + " is not a valid field for saving")
            field = getattr(state.fields, name)
            self.fields_save_metadata[name] = {
                "components": 1 if field.ndim == 1 else int(field.shape[1]),
                "dtype": str(field.dtype)}
        self._dirs_made = False
        self.state = state
        print_log("\nSetting up I/O operator done!", rank, verbose)
        print_log("-" * 80, rank, verbose)

    def _prepare_dirs(self, state):
        if self._dirs_made:
            return
        if state.domain.mpi_size == 1:
            self.field_save_path = os.path.join(self.root_dir, "output",
                                                "fields")
        else:
            self.field_save_path = os.path.join(
                self.root_dir, "procs", "proc_" + str(state.domain.mpi_rank))
        os.makedirs(self.field_save_path, exist_ok=True)
        self.dump_metadata(state)
        self._dirs_made = True

    def dump_metadata(self, state):
        from .solver import __version__
        c = state.control
        self.global_metadata = {
            "pylabolt": {"version": __version__,
                         "solver": self.model.solver_name},
            "control": {"end_time": c.end_time, "start_time": c.start_time,
                        "save_interval": c.save_interval,
                        "checkpoint_interval": c.checkpoint_interval},
            "mesh": {"size": int(state.mesh.grid_global_size),
                     "shape": (int(state.mesh.grid_global_shape[0]),
                               int(state.mesh.grid_global_shape[1]))},
            "decomposition": {"nx": int(state.domain.no_of_procs_x),
                              "ny": int(state.domain.no_of_procs_y)},
            "fields_saved": self.fields_save_metadata}
        if state.domain.mpi_rank == 0:
            with open(os.path.join(self.root_dir, "metadata.json"), "w") as f:
                json.dump(self.global_metadata, f, indent=4)
        if state.domain.mpi_size > 1:
            d = state.domain
            self.rank_metadata = {
                "rank": int(d.mpi_rank),
                "processor_ij": (int(d.i_proc), int(d.j_proc)),
                "domain_size": int(d.inner_size),
                "domain_shape": (int(d.inner_shape[0]), int(d.inner_shape[1])),
                "offset": (int(d.offset[0]), int(d.offset[1]))}
            with open(os.path.join(self.field_save_path,
                                   "rank_metadata.json"), "w") as f:
                json.dump(self.rank_metadata, f, indent=4)

    def set_backend(self, state, backend, plb):
        self.plb = plb

    # -- histories: utils/io_operator.py:236-357 ------------------------------
    def _history_path(self, name):
        return os.path.join(self.root_dir, "output", "histories", name + ".dat")

    def _keep_history(self, name):
        return (getattr(self, "_histories_resume", False) and
                os.path.exists(self._history_path(name)))

    def setup_write_histories(self, state):
        """Headers of output/histories/<name>.dat, one file per obstacle and
        per wall boundary element, written by rank 0."""
        self._histories_ready = True
        wanted = (state.obstacle.write_obstacle_data or
                  state.boundary.write_boundary_data)
        if not wanted or state.domain.mpi_rank != 0:
            return
        os.makedirs(os.path.join(self.root_dir, "output", "histories"),
                    exist_ok=True)
        if state.obstacle.write_obstacle_data:
            for body in state.obstacle.obstacles:
                if self._keep_history(body.name):
                    continue
                with open(self._history_path(body.name), "w") as out:
                    out.write(f"{'#':5} {'PyLaBolt obstacle history'}\n"
                              f"{'#':5} {'ID':8}: {body.id}\n"
                              f"{'#':5} {'Name':8}: {body.name}\n"
                              f"{'#':5} {'Type':8}: {body.type}\n"
                              f"{'#':5} {'Columns':8}:\n")
                    out.write(f"{'#':5}{'time':<21}" + "".join(
                        f"{col:<30}" for col in
                        ("pos_x", "pos_y", "alpha", "vel_x", "vel_y", "omega",
                         "force_x", "force_y", "torque")) + "\n")
        if state.boundary.write_boundary_data:
            for element in state.boundary.boundary_elements:
                if not element.wall or self._keep_history(element.name):
                    continue
                with open(self._history_path(element.name), "w") as out:
                    out.write(f"{'#':5} {'PyLaBolt boundary history'}\n"
                              f"{'#':5} {'Name':8}: {element.name}\n"
                              f"{'#':5} {'Columns':8}:\n")
                    out.write(f"{'#':5}{'time':<21}{'force_x':<30}"
                              f"{'force_y':<30}\n")

    def write_histories(self, state, time_step):
        if not getattr(self, "_histories_ready", False):
            self.setup_write_histories(state)
        if state.domain.mpi_rank != 0:
            return
        ob = state.obstacle
        if (ob.write_obstacle_data and ob.write_interval is not None and
                time_step % ob.write_interval == 0):
            for body in ob.obstacles:
                values = (body.center[0], body.center[1],
                          body.inclination_angle, body.linear_velocity[0],
                          body.linear_velocity[1], body.angular_velocity,
                          body.force[0], body.force[1], body.torque)
                with open(self._history_path(body.name), "a") as out:
                    out.write(f"{time_step:<24}" + "".join(
                        f"{float(v):<30.16e}" for v in values) + "\n")
        bd = state.boundary
        if (bd.write_boundary_data and bd.write_interval is not None and
                time_step % bd.write_interval == 0):
            for element in bd.boundary_elements:
                if not element.wall:
                    continue
                with open(self._history_path(element.name), "a") as out:
                    out.write(f"{time_step:<24}{element.force[0]:<30.16e}"
                              f"{element.force[1]:<30.16e}\n")

    # -- checkpoints: parsed but never implemented upstream (control.py:35) ----
    def _checkpoint_path(self, state, time_step):
        if state.domain.mpi_size == 1:
            base = os.path.join(self.root_dir, "output", "checkpoints")
        else:
            base = os.path.join(self.root_dir, "procs",
                                "proc_" + str(state.domain.mpi_rank))
        return os.path.join(base, "checkpoint_t_" + str(time_step) + ".npz")

    def write_checkpoint(self, state, time_step):
        """Every checkpoint_interval steps: this rank's pop_fluid_new in the
        reference layout (all a restart needs: rho and u are its moments)."""
        interval = state.control.checkpoint_interval
        if interval is None or time_step % interval != 0:
            return
        path = self._checkpoint_path(state, time_step)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez(path, pop_fluid_new=self.plb.download(capi.POP),
                 time_step=np.int64(time_step),
                 shape=np.asarray(state.domain.shape, dtype=np.int64))

    def has_checkpoint(self, state, time_step):
        return bool(time_step) and os.path.exists(
            self._checkpoint_path(state, time_step))

    def resume_histories(self, state):
        """A restarted run continues the history files of the run it resumes
        (rows are appended; headers only for files that do not exist yet)."""
        self._histories_resume = True

    def load_checkpoint(self, state, time_step):
        """start_time > 0 resumes from the checkpoint of that step, if one
        exists.  Returns True when populations were restored."""
        if not time_step:
            return False
        path = self._checkpoint_path(state, time_step)
        if not os.path.exists(path):
            return False
        data = np.load(path)
        if not np.array_equal(data["shape"], state.domain.shape):
            raise ValueError("checkpoint " + path + " was written for a "
                             "different decomposition")
        self.plb.upload(capi.POP, data["pop_fluid_new"])
        return True

    def _pinned(self, name, shape):
        """Page-locked download buffers, allocated (and faulted in) once and
        reused by every output step."""
        if not hasattr(self, "_pinned_buffers"):
            self._pinned_buffers = {}
        buf = self._pinned_buffers.get(name)
        if buf is None:
            if not hasattr(self.plb, "pinned"):
                return None             # host stand-in of the I/O tests
            buf = self.plb.pinned(shape)
            buf.array[...] = 0.0
            self._pinned_buffers[name] = buf
        return buf.array

    def gather_fields(self, state):
        """The arrays of one output file, ghost ring stripped.  rho and u are
        views of the reused pinned buffers (valid until the next call); the
        flag fields of static bodies never change and are stripped once."""
        out = {}
        n_inner = int(state.domain.inner_size)
        if not hasattr(self, "_static_fields"):
            self._static_fields = {}
        for name in self.fields_list:
            if name == "density":
                out[name] = self.plb.download(
                    capi.DENSITY_INNER, out=self._pinned(name, (n_inner,)))
            elif name == "velocity":
                out[name] = self.plb.download(
                    capi.VELOCITY_INNER, out=self._pinned(name, (n_inner, 2)))
            else:
                if name not in self._static_fields:
                    self._static_fields[name] = strip_ghost(
                        getattr(state.fields, name), state.domain.shape)
                out[name] = self._static_fields[name]
        return out

    def close(self):
        for buf in getattr(self, "_pinned_buffers", {}).values():
            buf.free()
        self._pinned_buffers = {}

    def write_fields(self, state, backend, time_step):
        interval = state.control.save_interval
        if interval is None or time_step % interval != 0:
            return
        self._prepare_dirs(state)
        np.savez(os.path.join(self.field_save_path,
                              "t_" + str(time_step) + ".npz"),
                 **self.gather_fields(state))
