#!/usr/bin/env python
"""Generates tests/golden/mrt_matrices.npz from the REFERENCE's own code.

The reference declares the MRT moment matrix and relaxation rates
(pylabolt/base/collision_operator.py:147-163, setup_MRT_params) but has no MRT
kernel, and setup_MRT_params itself cannot finish (np.matmul of the 9 x 9
inverse with the 1-D rate vector, then np.linalg.inv of a vector, :164-165;
its only call site passes no `state`, :93).  Everything it computes BEFORE
that point -- M, inv_M = np.linalg.inv(M), S -- is taken here by running the
reference's function on stand-in objects and catching the failure.

    python tests/golden/make_mrt_matrices.py      (needs /root/reference)
"""
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402

make_golden.install_shims()
from pylabolt.base.collision_operator import CollisionOperator  # noqa: E402

out = {}
for kin_visc in (0.1, 0.08, 0.02):
    lattice_inv_cs_2 = 1.0 / (np.float64(1 / np.sqrt(3)) * np.float64(1 / np.sqrt(3)))
    tau = kin_visc * lattice_inv_cs_2 + 0.5            # :89-91
    this = SimpleNamespace(omega_fluid=1 / tau)
    state = SimpleNamespace(control=SimpleNamespace(precision=np.float64))
    failure = None
    try:
        CollisionOperator.setup_MRT_params(this, state)
    except Exception as e:                               # upstream defect, see above
        failure = type(e).__name__
    tag = f"nu{kin_visc}"
    out[tag + "_M"] = this.M
    out[tag + "_inv_M"] = this.inv_M
    out[tag + "_S"] = this.S
    out[tag + "_omega"] = np.float64(this.omega_fluid)
    print(tag, "omega", this.omega_fluid, "upstream setup ended with", failure)
np.savez_compressed(os.path.join(HERE, "mrt_matrices.npz"), **out)
print("wrote tests/golden/mrt_matrices.npz")
