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

    def gather_fields(self, state):
        """The arrays of one output file, ghost ring stripped."""
        out = {}
        for name in self.fields_list:
            if name == "density":
                out[name] = self.plb.download(capi.DENSITY_INNER)
            elif name == "velocity":
                out[name] = self.plb.download(capi.VELOCITY_INNER)
            else:
                out[name] = strip_ghost(getattr(state.fields, name),
                                        state.domain.shape)
        return out

    def write_fields(self, state, backend, time_step):
        interval = state.control.save_interval
        if interval is None or time_step % interval != 0:
            return
        self._prepare_dirs(state)
        np.savez(os.path.join(self.field_save_path,
                              "t_" + str(time_step) + ".npz"),
                 **self.gather_fields(state))
