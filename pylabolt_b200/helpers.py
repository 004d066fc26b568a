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


def _format_residue(value):
    """One logged quantity: a scalar, a one-component array or a vector."""
    parts = np.atleast_1d(np.asarray(value, dtype=np.float64)).ravel()
    shown = ["%.5e" % v for v in parts]
    return shown[0] if len(shown) == 1 else "(" + ", ".join(shown) + ")"


class SimulationStatusLogger:
    """The status line of the reference (pylabolt/utils/helpers.py:29-56):
    ``time: <step> | <name>: <value> | ...`` on rank 0, on the steps that are
    multiples of ``control.std_out_interval``.  The column widths are part of
    the output contract (log scrapers split on them)."""

    TIME_LABEL = "%-5s %-10d"
    FIELD = "%-5s: %s"

    def __init__(self, mpi_rank, verbose=True):
        self.verbose = verbose

    def due(self, state, time_step):
        every = state.control.std_out_interval
        return every is not None and time_step % every == 0

    def log_data(self, state, time_step, **values):
        if not self.due(state, time_step):
            return
        line = [self.TIME_LABEL % ("time:", time_step)]
        line += [self.FIELD % (name, _format_residue(value))
                 for name, value in values.items()]
        print_log(" | ".join(line), state.domain.mpi_rank, self.verbose)
