#!/usr/bin/env python
"""Runs the REFERENCE's own unit tests (tests/unit of the reference checkout)
twice -- against the reference's containers and against the host mirror of
pylabolt_b200 -- and records the outcome of every test.

    python tools/reference_unit_conformance.py [--write]

The host mirror keeps the reference's class / function names and signatures
(Control, Mesh, Lattice, Fields, init_fields, read_dict, Boundary, Obstacle),
so the reference's tests can drive it unchanged: a throw-away alias package
maps ``pylabolt.base.<module>`` onto ``pylabolt_b200``.  The reference arm
needs one stand-in (mpi4py is not installed here).  Nothing is copied from the
reference; its test files are executed where they lie.

What the result means: every test the reference passes on its own containers
must pass on the mirror.  (About half of the reference's tests fail on the
reference itself -- its tests drifted from its case-file schema, SURVEY.md
section 4 -- and fail identically on the mirror, which raises the same
messages; test_domain.py / test_mpi_operator.py cannot even be imported
against the reference snapshot.  The mirror runs test_domain.py -- its 39
decomposition tests all pass, "reference: missing" in the table -- and leaves
out test_mpi_operator.py: the halo exchange lives on the device here.)

Second pass, "drift repaired": most of those failures have one cause -- the
reference's test dictionaries lack the ``wall`` key its Boundary has since
required.  A pytest plugin (generated here, applied to BOTH arms) fills in
``wall: False`` before Boundary.__init__ runs.  Then the reference passes 68
tests, among them the boundary node / direction lists for one and several ranks
and the periodic pairs (TestBoundaryNodeAllocation), and the mirror must pass
every one of them except the phase-field boundary sections
(TestPhaseBoundaryDicts: phaseFieldLB is outside the b200 build).

--write stores the table in tests/golden/reference_unit_conformance.json
(the CPU test suite checks the committed table, and re-runs the mirror arm
wherever the reference checkout is present).
"""
import json
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("PLB_REFERENCE", "/root/reference")
UNIT = os.path.join(REFERENCE, "tests", "unit")
SKIPPED_MODULES = ["test_domain.py", "test_mpi_operator.py"]
TABLE = os.path.join(REPO, "tests", "golden", "reference_unit_conformance.json")

# pylabolt.<...> module of the reference -> names taken from pylabolt_b200
ALIASES = {
    "pylabolt/base/control.py": "from pylabolt_b200.state import Control",
    "pylabolt/base/mesh.py": "from pylabolt_b200.state import Mesh",
    "pylabolt/base/lattice.py": "from pylabolt_b200.state import Lattice",
    "pylabolt/base/fields.py": "from pylabolt_b200.state import Fields",
    "pylabolt/base/init_fields.py":
        "from pylabolt_b200.state import (init_fields, read_dict, set_field_scalar,\n"
        "    set_field_vector, local_to_global, global_to_local)",
    "pylabolt/base/boundary.py":
        "from pylabolt_b200.boundary import Boundary, BoundaryElement",
    "pylabolt/base/obstacle.py":
        "from pylabolt_b200.obstacle import Obstacle, Circle, Ellipse",
    "pylabolt/parallel/domain.py":
        "from pylabolt_b200.state import Domain, local_to_global, global_to_local",
}

MPI_STAND_IN = '''
class _Comm:
    def Get_rank(self): return 0
    def Get_size(self): return 1
    def Barrier(self): pass
    def Abort(self, code=1): raise RuntimeError("comm.Abort()")
    def Sendrecv(self, sendbuf, dest, sendtag, recvbuf, source, recvtag):
        recvbuf[...] = sendbuf
    def Allreduce(self, local, out, op=None): out[...] = local
COMM_WORLD = _Comm()
SUM = "sum"
def Init(): pass
def Finalize(): pass
'''


DRIFT_REPAIR_PLUGIN = '''
import pytest


@pytest.fixture(autouse=True)
def _wall_key_default(monkeypatch):
    import pylabolt.base.boundary as module
    original = module.Boundary.__init__

    def init(self, simulation, *args, **kwargs):
        boundary_dict = getattr(simulation, "boundary_dict", None)
        if isinstance(boundary_dict, dict):
            for name, user in boundary_dict.items():
                if name != "options" and isinstance(user, dict):
                    user.setdefault("wall", False)
        return original(self, simulation, *args, **kwargs)
    monkeypatch.setattr(module.Boundary, "__init__", init)
'''
OUT_OF_SCOPE = "TestPhaseBoundaryDicts"      # phase-field boundary sections


def write_tree(root, files):
    for rel, text in files.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(text + "\n")
    for base, _, _ in os.walk(root):
        init = os.path.join(base, "__init__.py")
        if base != root and not os.path.exists(init):
            open(init, "w").close()


def run_arm(python_path, repair=False, skipped=tuple(SKIPPED_MODULES)):
    cmd = [sys.executable, "-m", "pytest", UNIT, "-q", "-p", "no:cacheprovider",
           "-rA", "--tb=no"]
    if repair:
        cmd += ["-p", "plb_drift_repair"]
    for name in skipped:
        cmd.append("--ignore=" + os.path.join(UNIT, name))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(python_path + [UNIT]),
               PYTHONDONTWRITEBYTECODE="1")
    with tempfile.TemporaryDirectory() as cwd:
        out = subprocess.run(cmd, env=env, cwd=cwd, capture_output=True,
                             text=True).stdout
    outcomes = {}
    for line in out.splitlines():
        m = re.match(r"^(PASSED|FAILED|ERROR) (\S+)", line)
        if m:
            test = m.group(2)
            outcomes[test[test.index("test_"):]] = m.group(1).lower()
    return outcomes


def mirror_arm(repair=False):
    with tempfile.TemporaryDirectory() as shim:
        write_tree(shim, dict(ALIASES, **{"plb_drift_repair.py": DRIFT_REPAIR_PLUGIN}))
        # test_domain.py needs local_to_global / global_to_local next to
        # Domain, which the mirror has and the reference snapshot has not
        return run_arm([shim, REPO], repair, skipped=SKIPPED_MODULES[1:])


def reference_arm(repair=False):
    with tempfile.TemporaryDirectory() as shim:
        write_tree(shim, {"mpi4py/__init__.py": "def rc(**kw): pass\nfrom . import MPI",
                          "mpi4py/MPI.py": MPI_STAND_IN,
                          "plb_drift_repair.py": DRIFT_REPAIR_PLUGIN})
        return run_arm([shim, REFERENCE], repair)


def compare(repair):
    ref, ours = reference_arm(repair), mirror_arm(repair)
    table = {name: {"reference": ref.get(name, "missing"),
                    "mirror": ours.get(name, "missing")}
             for name in sorted(set(ref) | set(ours))}
    count = {}
    for row in table.values():
        key = row["reference"] + " -> " + row["mirror"]
        count[key] = count.get(key, 0) + 1
    print("drift repaired:" if repair else "as shipped:")
    for key in sorted(count):
        print(f"{count[key]:4d}  reference {key} (mirror)")
    bad = [n for n, row in table.items()
           if row["reference"] == "passed" and row["mirror"] != "passed"
           and OUT_OF_SCOPE not in n]
    for name in bad:
        print("NOT CONFORMING:", name)
    return table, bad


def main():
    if not os.path.isdir(UNIT):
        raise SystemExit("reference checkout not found at " + REFERENCE)
    table, bad = compare(False)
    repaired, bad_repaired = compare(True)
    if "--write" in sys.argv:
        with open(TABLE, "w") as f:
            json.dump({"skipped_modules": SKIPPED_MODULES,
                       "out_of_scope": OUT_OF_SCOPE, "tests": table,
                       "tests_drift_repaired": repaired}, f, indent=1, sort_keys=True)
            f.write("\n")
        print("wrote", os.path.relpath(TABLE, REPO))
    return 1 if bad or bad_repaired else 0


if __name__ == "__main__":
    sys.exit(main())
