"""pylabolt_b200 -- B200-native back end of PyLaBolt's fluidLB time step.

Hand-written sm_100a CUDA (pylabolt_b200/csrc) behind a C ABI
(include/plb.h), driven by a Python host layer that mirrors the reference's
case-file schema and Solver seam.  See DESIGN.md.
"""
__version__ = "1.0.0.dev0+b200"
