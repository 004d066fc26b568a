"""Small host utilities with the reference's names and behaviour
(pylabolt/utils/helpers.py:6-56)."""
import importlib
import os
import sys

import numpy as np


def print_log(mssg, mpi_rank, verbose):
    """Rank-0 print (pylabolt/utils/helpers.py:6-8)."""
    if verbose and mpi_rank == 0:
        print(mssg, flush=True)


def load_simulation(comm, mpi_rank):
    """Imports the user's case file ``simulation.py`` from the current working
    directory (pylabolt/utils/helpers.py:11-27)."""
    try:
        cwd = os.getcwd()
        if cwd not in sys.path:
            sys.path.append(cwd)
        if not os.path.exists(os.path.join(cwd, "simulation.py")):
            raise ImportError(
                "Missing simulation.py file in current working directory")
        sys.modules.pop("simulation", None)
        return importlib.import_module("simulation")
    except Exception as e:
        print_log("-" * 80, mpi_rank, True)
        print_log("FATAL ERROR!", mpi_rank, True)
        print_log(str(e), mpi_rank, True)
        comm.Abort()
        raise


class SimulationStatusLogger:
    """``time | res_density | res_velocity`` line every std_out_interval steps
    (pylabolt/utils/helpers.py:29-56)."""

    def __init__(self, mpi_rank, verbose=True):
        self.verbose = verbose

    def log_data(self, state, time_step, **values):
        interval = state.control.std_out_interval
        if interval is None or time_step % interval != 0:
            return
        parts = [f"{'time:':<5} {time_step:<10}"]
        for key, value in values.items():
            if np.isscalar(value):
                text = f"{value:.5e}"
            elif len(value) == 1:
                text = f"{value[0]:.5e}"
            else:
                text = "(" + ", ".join(f"{v:.5e}" for v in value) + ")"
            parts.append(f"{key:<5}: {text}")
        print_log(" | ".join(parts), state.domain.mpi_rank, self.verbose)
