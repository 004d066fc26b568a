"""TEST INFRASTRUCTURE: runs a script (bench.py) against the emulated libplb.

    PLB_LIB=<libplb_emu.so> python tests/emu/run_emulated.py script.py [args]

The product loader (pylabolt_b200/capi.py) refuses an emulated build; this
wrapper is the one place outside pytest that lifts the refusal, so that the
script's own code stays free of any CPU path.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from pylabolt_b200 import capi  # noqa: E402

capi._accept_emulated_build = True
script = sys.argv[1]
sys.argv = sys.argv[1:]
runpy.run_path(script, run_name="__main__")
