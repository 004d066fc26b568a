"""Field output contract: the b200 InputOutputOperator writes the same
metadata.json structure and the same .npz arrays (names, shapes, dtypes,
values) as the reference's (utils/io_operator.py:96-190); the golden copy was
written by the reference itself (tests/golden/make_golden.py, cylinder case).
The device is replaced by a stand-in that serves rho / u from host arrays, so
this runs without a GPU."""
import json
import os

import numpy as np

import cases
from pylabolt_b200 import capi
from pylabolt_b200.comm import SingleComm
from pylabolt_b200.io_operator import InputOutputOperator, strip_ghost
from pylabolt_b200.operators import FluidLB
from pylabolt_b200.state import State


class HostPlb:
    """plb_download of the inner region, served from the host fields."""

    def __init__(self, state):
        self.state = state

    def download(self, field, out=None):
        f, shape = self.state.fields, self.state.domain.shape
        if field == capi.DENSITY_INNER:
            return strip_ghost(f.density, shape)
        if field == capi.VELOCITY_INNER:
            return strip_ghost(f.velocity, shape)
        raise AssertionError(field)


def test_npz_and_metadata_match_the_reference(golden_dir, tmp_path):
    data = np.load(os.path.join(golden_dir, "cylinder.npz"))
    factory, kwargs, _ = cases.GOLDEN_CASES["cylinder"]
    sim = factory(**kwargs)
    sim.control_dict["save_interval"] = 5
    st = State(sim, SingleComm(), 0, verbose=False)
    io = InputOutputOperator(FluidLB(), st, None, SingleComm(), verbose=False,
                             root_dir=str(tmp_path))
    io.set_backend(st, None, HostPlb(st))
    io.write_fields(st, None, 3)                 # not a save step
    assert not (tmp_path / "output").exists()
    io.write_fields(st, None, 0)
    saved = np.load(tmp_path / "output" / "fields" / "t_0.npz")
    want = {k[3:]: data[k] for k in data.files
            if k.startswith("io_") and k != "io_metadata_json"}
    assert list(saved.files) == list(want)      # same order as save_fields
    for key, ref in want.items():
        assert saved[key].shape == ref.shape and saved[key].dtype == ref.dtype
        assert np.array_equal(saved[key], ref), key

    ours = json.load(open(tmp_path / "metadata.json"))
    theirs = json.loads(str(data["io_metadata_json"]))
    assert ours.keys() == theirs.keys()
    for section in ("mesh", "decomposition", "fields_saved"):
        assert ours[section] == theirs[section]
    assert ours["control"].keys() == theirs["control"].keys()
    assert ours["pylabolt"]["solver"] == theirs["pylabolt"]["solver"] == "fluidLB"


def test_multi_rank_layout(tmp_path):
    """procs/proc_<r>/ + rank_metadata.json (io_operator.py:82-95, 135-156)."""
    from test_decomposition import DummyComm
    sim = cases.cavity(37, 29)
    sim.control_dict["save_interval"] = 1
    sim.decompose_dict = {"nx": 2, "ny": 1}
    for rank in range(2):
        st = State(sim, DummyComm(rank, 2), rank, verbose=False)
        io = InputOutputOperator(FluidLB(), st, None, DummyComm(rank, 2),
                                 verbose=False, root_dir=str(tmp_path))
        io.set_backend(st, None, HostPlb(st))
        io.write_fields(st, None, 0)
        meta = json.load(open(tmp_path / "procs" / f"proc_{rank}" /
                              "rank_metadata.json"))
        assert meta["rank"] == rank and meta["processor_ij"] == [rank, 0]
        assert meta["offset"] == [19 * rank, 0]
        assert meta["domain_shape"] == [19 if rank == 0 else 18, 29]
        saved = np.load(tmp_path / "procs" / f"proc_{rank}" / "t_0.npz")
        assert saved["density"].shape == (meta["domain_size"],)
    assert json.load(open(tmp_path / "metadata.json"))["decomposition"] == \
        {"nx": 2, "ny": 1}
