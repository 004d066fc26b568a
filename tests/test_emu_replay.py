"""Replays GPU test modules against the emulated library (test infrastructure,
tests/emu/README.md): the `-m gpu` tests only talk to the C ABI, so the same
assertions run on the CPU when PLB_LIB points at libplb_emu.so.  Used here for
the Solver life cycle (run() output contract, forces, residues, checkpoint /
restart) with steps grouped two and three per pass, with the emulated device
allocations flush against a guard page at their lower end (PLB_EMU_GUARD=lo;
the default places the guard at the upper end), so that an out-of-range row or
column access of a kernel is a crash, not a silent read of a neighbour.
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
import build_emu  # noqa: E402


SOLVER = "solver_run or residues or mass"
ORACLE = ("(against_oracle and (cavity_101 or wide_row or odd_sizes or "
          "columns_periodic) and True)")


@pytest.mark.parametrize("depth,guard,select", [("2", "lo", SOLVER),
                                                ("3", "lo", SOLVER + " or " + ORACLE)])
def test_solver_life_cycle_on_the_emulator(depth, guard, select):
    lib = build_emu.build()
    env = dict(os.environ, PLB_LIB=lib, PLB_FUSE="2", PLB_FUSE_DEPTH=depth,
               PLB_EMU_GUARD=guard, PLB_EMU_TESTING="1")
    proc = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
         os.path.join(HERE, "test_gpu_solver_run.py"),
         os.path.join(HERE, "test_gpu_parity.py"), "-k", select],
        capture_output=True, text=True, timeout=1500, env=env, cwd=os.path.dirname(HERE))
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-2000:]
    assert " passed" in proc.stdout


def test_smoke_entry_point_on_the_emulator():
    """__graft_entry__.smoke() -- what the driver runs on cuda:0 before the
    bench -- replayed against the emulated library: its own bookkeeping (how
    many of its steps the several-steps-per-pass kernel took) must agree with
    the shipped defaults."""
    lib = build_emu.build()
    env = dict(os.environ, PLB_LIB=lib, PLB_EMU_TESTING="1")
    env.pop("PLB_FUSE", None)
    env.pop("PLB_FUSE_DEPTH", None)
    proc = subprocess.run(
        [sys.executable, "-c",
         "from pylabolt_b200 import capi; capi._accept_emulated_build = True; "
         "import __graft_entry__ as g; g.smoke()"],
        capture_output=True, text=True, timeout=900, env=env, cwd=os.path.dirname(HERE))
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-3000:]
    assert "10 four-step" in proc.stdout
