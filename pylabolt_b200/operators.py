"""Collision and forcing configuration: the b200 mirror of the parsing half of
pylabolt/base/collision_operator.py and pylabolt/base/force_operator.py.

In the reference these operators also pick numba kernels by name
(``{model}_{equilibrium}_{forcing}``, collision_operator.py:448-461); here the
same three strings select a template instantiation of the fused CUDA kernel
inside libplb, so the operators only validate and hold parameters.
"""
import numpy as np

from .helpers import print_log


class FluidLB:
    """Model descriptor (pylabolt/solvers/fluidLB.py:22-60)."""

    def __init__(self):
        self.solver_name = "fluidLB"
        self.equilibrium_models = {"fluid": ["density_based_second_order"]}
        self.forcing_models = {"fluid": [None, "guo_linear",
                                         "guo_second_order"]}
        self.collision_models = {"fluid": ["BGK", "MRT"]}
        self.streaming_type = {"fluid": "scalar_based"}
        self.boundary_condition_type = {"fluid": "density_based"}
        self.compute_fields_config = {"type": "density_based",
                                      "moment_fields": ["density", "velocity"]}
        self.obstacle_kernels_type = "single_phase"
        self.residue_fields = ["density", "velocity"]
        self.save_fields = ["density", "velocity", "solid", "solid_id",
                            "solid_boundary", "fluid_boundary",
                            "surface_normals"]


def _abort(state, comm, error):
    rank = state.domain.mpi_rank
    print_log("-" * 80, rank, True)
    print_log("FATAL ERROR!", rank, True)
    print_log(str(error), rank, True)
    comm.Abort()


class CollisionOperator:
    """collision_dict.fluid -> model, equilibrium, forcing, tau, omega
    (base/collision_operator.py:44-134)."""

    def __init__(self, simulation, model, state, comm, verbose=True):
        rank = state.domain.mpi_rank
        try:
            print_log("-" * 80, rank, verbose)
            print_log("Setting up collision operator...\n", rank, verbose)
            if not hasattr(simulation, "collision_dict"):
                raise ValueError(
                    "collision_dict not found in simulation.py file")
            self.model = model
            self.collision_dict = simulation.collision_dict
            self.read_collision_dict(state, verbose)
            print_log("\nSetting up collision operator done!", rank, verbose)
            print_log("-" * 80, rank, verbose)
        except Exception as e:
            _abort(state, comm, e)
            raise

    def read_collision_dict(self, state, verbose=True):
        if "fluid" not in self.collision_dict:
            raise ValueError("fluid missing in collision_dict")
        fluid = self.collision_dict["fluid"]
        for key in ("model", "equilibrium", "forcing_model"):
            if key not in fluid:
                raise ValueError(key + " missing in fluid: collision_dict")
        self.collision_fluid = fluid["model"]
        self.equilibrium_fluid = fluid["equilibrium"]
        self.forcing_fluid = fluid["forcing_model"]
        if self.forcing_fluid == "None":
            self.forcing_fluid = None
        checks = (
            (self.collision_fluid, self.model.collision_models["fluid"],
             "Unsupported fluid collision model: "),
            (self.equilibrium_fluid, self.model.equilibrium_models["fluid"],
             "Unsupported fluid equilibrium model: "),
            (self.forcing_fluid, self.model.forcing_models["fluid"],
             "Unsupported fluid forcing model: "))
        for value, allowed, message in checks:
            if value not in allowed:
                raise ValueError(message + str(value) +
                                 "\nAvailable models: " + str(allowed))
        # base/collision_operator.py:89-91
        self.tau_fluid = state.transport.kin_visc * state.lattice.inv_cs_2 + 0.5
        self.omega_fluid = 1 / self.tau_fluid
        if self.collision_fluid == "MRT":
            self.setup_MRT_params(state)
            self.collision_params = (self.M, self.inv_M, self.S)
        else:
            self.collision_params = (self.omega_fluid,)
        rank = state.domain.mpi_rank
        print_log(f"{'Fluid collision model':<30}: {self.collision_fluid}",
                  rank, verbose)
        print_log(f"{'Fluid equilibrium model':<30}: {self.equilibrium_fluid}",
                  rank, verbose)
        print_log(f"{'Fluid forcing model':<30}: {str(self.forcing_fluid)}",
                  rank, verbose)

    def setup_MRT_params(self, state):
        """The Lallemand-Luo matrix and rates the reference declares
        (base/collision_operator.py:147-163).  Upstream's own use of them is
        broken (:93, :164-165); ours is SURVEY.md App. A.2.  ``mrt_rates`` in
        collision_dict.fluid may override S (nine rates)."""
        prec = state.control.precision
        self.M = np.array([
            [1, 1, 1, 1, 1, 1, 1, 1, 1],
            [-4, -1, -1, -1, -1, 2, 2, 2, 2],
            [4, -2, -2, -2, -2, 1, 1, 1, 1],
            [0, 1, 0, -1, 0, 1, -1, -1, 1],
            [0, -2, 0, 2, 0, 1, -1, -1, 1],
            [0, 0, 1, 0, -1, 1, 1, -1, -1],
            [0, 0, -2, 0, 2, 1, 1, -1, -1],
            [0, 1, -1, 1, -1, 0, 0, 0, 0],
            [0, 0, 0, 0, 0, 1, -1, 1, -1]], dtype=prec)
        self.inv_M = self.M.T / np.sum(self.M * self.M, axis=1)
        rates = self.collision_dict["fluid"].get("mrt_rates")
        if rates is None:
            rates = [1.0] * 7 + [self.omega_fluid] * 2
        if len(rates) != 9:
            raise ValueError("mrt_rates must list nine relaxation rates")
        self.S = np.array(rates, dtype=prec)

    @property
    def mrt_rates(self):
        if self.collision_fluid == "MRT":
            return self.S
        return np.array([1.0] * 7 + [self.omega_fluid] * 2)


class ForceOperator:
    """forcing_dict.gravity (base/force_operator.py:46-81): ignored, with a
    warning, when collision_dict's forcing_model is None."""

    def __init__(self, simulation, model, state, comm, collision_operator=None,
                 verbose=True):
        rank = state.domain.mpi_rank
        try:
            print_log("-" * 80, rank, verbose)
            print_log("Setting up forcing operator...\n", rank, verbose)
            if not hasattr(simulation, "forcing_dict"):
                raise ValueError("forcing_dict not found in simulation.py file")
            self.model = model
            self.collision_operator = collision_operator
            self.forcing_dict = simulation.forcing_dict
            self.gravity = np.zeros(2, dtype=state.control.precision)
            if "gravity" in self.forcing_dict:
                if collision_operator.forcing_fluid is None:
                    print_log("WARNING! gravity ignored in forcing dict" +
                              " as forcing is set to None in collision dict\n",
                              rank, verbose)
                else:
                    gravity = self.forcing_dict["gravity"]
                    if not isinstance(gravity, list) or len(gravity) != 2:
                        raise ValueError(
                            "gravity must be a list (gx, gy) in forcing dict")
                    self.gravity = np.array(gravity,
                                            dtype=state.control.precision)
                print_log(f"{'Gravity':<30}: {self.gravity}", rank, verbose)
            print_log("\nSetting up forcing operator done!", rank, verbose)
            print_log("-" * 80, rank, verbose)
        except Exception as e:
            _abort(state, comm, e)
            raise
