"""--reconstruct and --to_vtk (pylabolt_b200/postprocess.py) on output written
by the b200 InputOutputOperator: same files and semantics as
pylabolt/utils/reconstruct.py and pylabolt/utils/npz2vtk.py.  No GPU."""
import json

import numpy as np
import pytest

import cases
from pylabolt_b200.comm import SingleComm
from pylabolt_b200.io_operator import InputOutputOperator
from pylabolt_b200.operators import FluidLB
from pylabolt_b200.postprocess import convert_to_vtk, reconstruct_data
from pylabolt_b200.state import State
from test_decomposition import DummyComm
from test_io_output import HostPlb


def _case(n_ranks):
    sim = cases.cylinder(41, 23)
    sim.control_dict["save_interval"] = 2
    sim.control_dict["end_time"] = 4
    sim.decompose_dict = {"nx": n_ranks, "ny": 1}
    return sim


def _write(root, n_ranks, times):
    sim = _case(n_ranks)
    for rank in range(n_ranks):
        comm = DummyComm(rank, n_ranks) if n_ranks > 1 else SingleComm()
        st = State(sim, comm, rank, verbose=False)
        # make the fields depend on the global position so that a misplaced
        # block cannot go unnoticed
        from pylabolt_b200.state import global_coordinates
        i_glob, j_glob = global_coordinates(st.domain)
        st.fields.density[:] = 1.0 + 1e-3 * i_glob + 1e-6 * j_glob
        st.fields.velocity[:, 0] = 1e-3 * j_glob
        st.fields.velocity[:, 1] = -1e-3 * i_glob
        io = InputOutputOperator(FluidLB(), st, None, comm, verbose=False,
                                 root_dir=str(root))
        io.set_backend(st, None, HostPlb(st))
        for t in times:
            st.fields.density += 1.0
            io.write_fields(st, None, t)


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_reconstruct_equals_single_rank_output(tmp_path, n_ranks):
    single, multi = tmp_path / "single", tmp_path / "multi"
    single.mkdir()
    multi.mkdir()
    _write(single, 1, (0, 2, 4))
    _write(multi, n_ranks, (0, 2, 4))
    reconstruct_data("all", root_dir=str(multi), verbose=False)
    for t in (0, 2, 4):
        want = np.load(single / "output" / "fields" / f"t_{t}.npz")
        got = np.load(multi / "output" / "fields" / f"t_{t}.npz")
        assert list(got.files) == list(want.files)
        for name in want.files:
            assert got[name].dtype == want[name].dtype, name
            assert np.array_equal(got[name], want[name]), (t, name)


def test_reconstruct_single_time_and_errors(tmp_path):
    _write(tmp_path, 2, (0, 2))
    reconstruct_data("time", time_step=2, root_dir=str(tmp_path), verbose=False)
    assert (tmp_path / "output" / "fields" / "t_2.npz").exists()
    assert not (tmp_path / "output" / "fields" / "t_0.npz").exists()
    with pytest.raises(FileNotFoundError):
        reconstruct_data("time", time_step=3, root_dir=str(tmp_path),
                         verbose=False)
    with pytest.raises(ValueError):
        reconstruct_data("some", root_dir=str(tmp_path), verbose=False)
    (tmp_path / "procs" / "proc_1" / "rank_metadata.json").unlink()
    with pytest.raises(FileNotFoundError):
        reconstruct_data("all", root_dir=str(tmp_path), verbose=False)


def _parse_legacy_vtk(path):
    """Minimal reader of the legacy RECTILINEAR_GRID files written above."""
    raw = open(path, "rb").read()
    pos = 0

    def line():
        nonlocal pos
        end = raw.index(b"\n", pos)
        text = raw[pos:end].decode()
        pos = end + 1
        return text

    assert line() == "# vtk DataFile Version 4.2"
    line()
    binary = line() == "BINARY"
    assert line() == "DATASET RECTILINEAR_GRID"
    dims = tuple(int(v) for v in line().split()[1:])

    def values(count, type_name):
        nonlocal pos
        dtype = ">f8" if type_name == "double" else ">i4"
        if binary:
            nbytes = count * np.dtype(dtype).itemsize
            a = np.frombuffer(raw[pos:pos + nbytes], dtype=dtype)
            pos += nbytes + 1
            return a.astype(np.float64 if type_name == "double" else np.int64)
        out = []
        while len(out) < count:
            out.extend(line().split())
        return np.array(out, dtype=np.float64 if type_name == "double"
                        else np.int64)

    coords = {}
    for axis in "XYZ":
        head = line().split()
        assert head[0] == axis + "_COORDINATES"
        coords[axis] = values(int(head[1]), head[2])
    n = int(line().split()[1])
    n_arrays = int(line().split()[2])
    arrays = {}
    for _ in range(n_arrays):
        name, comp, tuples, type_name = line().split()
        assert int(tuples) == n
        arrays[name] = values(int(comp) * n, type_name).reshape(n, int(comp))
    return dims, coords, arrays


@pytest.mark.parametrize("binary", [False, True])
def test_vtk_file_holds_every_saved_field_in_point_order(tmp_path, binary):
    _write(tmp_path, 1, (0, 2))
    convert_to_vtk("all", root_dir=str(tmp_path), verbose=False, binary=binary)
    meta = json.load(open(tmp_path / "metadata.json"))
    nx, ny = meta["mesh"]["shape"]
    for t in (0, 2):
        saved = np.load(tmp_path / "output" / "fields" / f"t_{t}.npz")
        dims, coords, arrays = _parse_legacy_vtk(
            tmp_path / "output" / "vtk" / f"t_{t}.vtk")
        assert dims == (nx, ny, 1)
        assert np.array_equal(coords["X"], np.arange(nx))
        assert np.array_equal(coords["Y"], np.arange(ny))
        assert np.array_equal(coords["Z"], [0.0])
        assert list(arrays) == ["point_ID"] + list(meta["fields_saved"])
        # point p = j * nx + i (x fastest) holds node ind = i * ny + j
        ind = arrays["point_ID"][:, 0]
        p = np.arange(nx * ny)
        assert np.array_equal(ind, (p % nx) * ny + p // nx)
        for name in meta["fields_saved"]:
            want = saved[name].reshape(nx * ny, -1)[ind]
            assert np.array_equal(arrays[name], want.astype(np.float64)
                                  if want.dtype.kind == "f"
                                  else want.astype(np.int64)), name


def test_vtk_rejects_output_that_disagrees_with_metadata(tmp_path):
    _write(tmp_path, 1, (0,))
    path = tmp_path / "output" / "fields" / "t_0.npz"
    data = dict(np.load(path))
    data["density"] = data["density"][:-1]
    np.savez(path, **data)
    with pytest.raises(ValueError):
        convert_to_vtk("time", 0, root_dir=str(tmp_path), verbose=False)
    with pytest.raises(FileNotFoundError):
        convert_to_vtk("time", 7, root_dir=str(tmp_path), verbose=False)


def test_cli_accepts_the_readme_spelling():
    from pylabolt_b200.cli import build_parser
    args = build_parser().parse_args(["--toVTK", "all"])
    assert args.to_vtk == "all"
    args = build_parser().parse_args(["--reconstruct", "time", "-t", "40"])
    assert args.reconstruct == "time" and args.time == 40
