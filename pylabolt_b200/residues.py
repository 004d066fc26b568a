"""Residues for the status log: mirror of pylabolt/utils/residues.py.

res_phi = sqrt( sum (phi - phi_old)^2 / (sum phi_old^2 + eps) ) over fluid
nodes, per component, evaluated every std_out_interval steps; phi_old is
updated only then (residues.py:171-222, cpu/compute_residues_kernels.py:6-73).
The sums run on the device over the stored moments (plb_residue_sums, a
deterministic two-stage reduction) and are summed across ranks.
"""
import numpy as np

from .helpers import print_log


class ResidueOperator:
    def __init__(self, model, state, comm, verbose=True):
        rank = state.domain.mpi_rank
        print_log("-" * 80, rank, verbose)
        print_log("Setting up residue operator...\n", rank, verbose)
        self.model = model
        self.fields_list = model.residue_fields
        self.residues = {
            "res_density": np.zeros(1, dtype=state.control.precision),
            "res_velocity": np.zeros(2, dtype=state.control.precision)}
        self.plb = None
        print_log("\nSetting up residue operator done!", rank, verbose)
        print_log("-" * 80, rank, verbose)

    def set_backend(self, state, backend, plb):
        self.plb = plb

    def compute_residues(self, state, backend, comm, time_step):
        interval = state.control.std_out_interval
        if interval is None or time_step % interval != 0:
            return
        local = self.plb.residue_sums()
        total = np.zeros_like(local)
        comm.Allreduce(local, total)
        eps = state.control.float_min
        self.residues["res_density"][0] = np.sqrt(total[0] / (total[1] + eps))
        self.residues["res_velocity"][0] = np.sqrt(total[2] / (total[3] + eps))
        self.residues["res_velocity"][1] = np.sqrt(total[4] / (total[5] + eps))
