"""The Solver life cycle on the GPU: run() output contract, forces, residues
log, checkpoint / restart -- what a user of `pylabolt --solver fluidLB` sees."""
import json
import os

import numpy as np
import pytest

import cases
from pylabolt_b200 import capi
from pylabolt_b200.comm import SingleComm
from pylabolt_b200.io_operator import strip_ghost
from pylabolt_b200.solver import Solver
from test_gpu_parity import make_solver, oracle_for

pytestmark = pytest.mark.gpu


def test_run_writes_reference_fields_at_reference_times(tmp_path, monkeypatch):
    """rho, u saved at loop index t are phases 2-4 of step t, i.e. moments of
    the lattice after t-1 streams; t = start_time holds the raw initial
    fields (solvers/fluidLB.py:339-343, 376-380)."""
    monkeypatch.chdir(tmp_path)
    sim = cases.cylinder(end_time=12)
    sim.control_dict["save_interval"] = 4
    sim.control_dict["std_out_interval"] = 6
    s = Solver(SingleComm(), "b200", simulation=sim, strict=True, verbose=False)
    s.set_backend()
    s.compile()
    orc = oracle_for(s)
    s.run()
    shape = s.state.domain.shape
    t0 = np.load("output/fields/t_0.npz")
    assert np.array_equal(t0["density"], strip_ghost(s.state.fields.density, shape))
    done = 0
    for t in (4, 8, 12):
        orc.step(t - done)
        done = t
        saved = np.load(f"output/fields/t_{t}.npz")
        assert np.array_equal(saved["density"], strip_ghost(orc.density, shape))
        assert np.array_equal(saved["velocity"], strip_ghost(orc.velocity, shape))
        assert np.array_equal(saved["solid"], strip_ghost(s.state.fields.solid, shape))
    assert sorted(os.listdir("output/fields")) == ["t_0.npz", "t_12.npz",
                                                   "t_4.npz", "t_8.npz"]
    meta = json.load(open("metadata.json"))
    assert meta["mesh"]["shape"] == [60, 31]
    # residues of the last logged step against the oracle's restatement
    orc2 = oracle_for(s)
    orc2.step(6)
    orc2.residues()
    orc2.step(6)
    want = orc2.residues()
    got = s.residue_operator.residues
    assert np.allclose([got["res_density"][0], *got["res_velocity"]], want,
                       rtol=1e-10)
    s.close()


@pytest.mark.parametrize("name", ["cavity", "cylinder", "cylinder_spin",
                                  "inflow_ellipse", "periodic_box"])
def test_forces_match_reference_kernels(golden_dir, name):
    factory, kwargs, record = cases.GOLDEN_CASES[name]
    data = np.load(os.path.join(golden_dir, name + ".npz"))
    sim = factory(**kwargs)
    sim.boundary_dict["options"] = {"compute_force": True,
                                    "write_boundary_data": {"interval": 1}}
    sim.obstacle_dict["options"] = dict(sim.obstacle_dict["options"],
                                        compute_force_torque=True)
    s = make_solver(sim, strict=True)
    # periodic elements own no links.  The golden forces come from the
    # reference's KERNEL called on every element; its operator (and ours)
    # only fills the rows of wall elements (boundary_operator.py:224-230).
    from pylabolt_b200.force_torque import MomentumExchange
    elements = s.state.boundary.boundary_elements
    real = [n for n, el in enumerate(elements) if el.type_fluid != "periodic"]
    not_wall = [n for n, el in enumerate(elements) if not el.wall]
    s.compute_forces(s.momentum.initial_exchange(s.state.lattice))
    assert not s.state.boundary.global_force[not_wall].any()
    assert not s.state.boundary.local_force[not_wall].any()
    s.momentum = MomentumExchange(s.state, s.plb.link_nodes(),
                                  every_element=True)
    try:
        wall, body = s.compute_forces(s.momentum.initial_exchange(s.state.lattice))
        assert np.abs(wall[real] - data["wall_force_0"][real]).max(initial=0) <= 1e-12
        done = 0
        for step in record:
            s.advance(step - done, record_links_last=True)
            done = step
            wall, body = s.compute_forces()
            want_wall = data[f"wall_force_{step}"][real]
            want_body = data[f"body_force_{step}"]
            assert np.abs(wall[real] - want_wall).max(initial=0) <= \
                1e-12 * max(1.0, np.abs(want_wall).max(initial=0))
            if want_body.size:
                assert np.abs(body - want_body).max() <= \
                    1e-12 * max(1.0, np.abs(want_body).max())
                assert np.array_equal(s.state.obstacle.obstacles[0].force, body[0, :2])
    finally:
        s.close()


def test_checkpoint_restart_is_bit_identical(tmp_path, monkeypatch):
    """A run split at a checkpoint leaves the same populations, the same
    output/fields files and the same history rows as the uninterrupted run:
    the restart neither rewrites t_<start>.npz with the case file's initial
    fields nor truncates the histories."""
    def solver(start, end, ckpt):
        sim = cases.periodic_box(end_time=end)
        sim.control_dict["start_time"] = start
        sim.control_dict["checkpoint_interval"] = ckpt
        sim.control_dict["save_interval"] = 5
        sim.obstacle_dict["options"] = dict(
            sim.obstacle_dict["options"], compute_force_torque=True,
            write_obstacle_data={"interval": 5})
        s = Solver(SingleComm(), "b200", simulation=sim, verbose=False)
        s.set_backend()
        s.compile()
        return s

    def outputs():
        fields = {name: dict(np.load(os.path.join("output/fields", name)))
                  for name in sorted(os.listdir("output/fields"))}
        hist = {name: open(os.path.join("output/histories", name)).read()
                for name in sorted(os.listdir("output/histories"))}
        return fields, hist

    (tmp_path / "whole").mkdir()
    monkeypatch.chdir(tmp_path / "whole")
    whole = solver(0, 30, None)
    whole.run()
    want = whole.plb.download(capi.POP)
    whole.close()
    want_fields, want_hist = outputs()
    assert "t_15.npz" in want_fields and want_hist

    (tmp_path / "split").mkdir()
    monkeypatch.chdir(tmp_path / "split")
    first = solver(0, 15, 15)
    first.run()
    first.close()
    assert os.path.exists("output/checkpoints/checkpoint_t_15.npz")
    second = solver(15, 30, None)
    second.run()
    got = second.plb.download(capi.POP)
    second.close()
    assert np.array_equal(got, want)
    got_fields, got_hist = outputs()
    assert sorted(got_fields) == sorted(want_fields)
    for name, arrays in want_fields.items():
        for key, value in arrays.items():
            assert np.array_equal(got_fields[name][key], value), (name, key)
    assert got_hist == want_hist


def test_pop_upload_download_round_trip():
    s = make_solver(cases.cylinder(), strict=False)
    try:
        s.advance(7)
        pop = s.plb.download(capi.POP)
        s.plb.upload(capi.POP, pop)
        assert np.array_equal(s.plb.download(capi.POP), pop)
        ghost = s.state.fields.ghost_node
        assert not pop[ghost].any() and not pop[s.state.fields.solid].any()
    finally:
        s.close()


def test_history_files_written_by_run(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    sim = cases.cylinder(end_time=6, spin=0.01)
    sim.obstacle_dict["options"]["write_obstacle_data"] = {"interval": 3}
    sim.boundary_dict["options"] = {"compute_force": True,
                                    "write_boundary_data": {"interval": 2}}
    s = Solver(SingleComm(), "b200", simulation=sim, verbose=False)
    s.set_backend()
    s.compile()
    orc = oracle_for(s)
    s.run()
    rows = [r for r in open("output/histories/cyl.dat").read().splitlines()
            if not r.startswith("#")]
    assert [int(r.split()[0]) for r in rows] == [0, 3, 6]
    orc.step(6)
    body = s.state.obstacle.obstacles[0]
    want = orc.obstacle_force_torque(s.state.fields.solid_id,
                                     s.state.fields.fluid_boundary,
                                     s.state.domain.offset,
                                     s.state.mesh.grid_global_shape,
                                     body.ref_point, body.id)
    got = [float(v) for v in rows[-1].split()[7:10]]
    assert np.allclose(got, want, rtol=1e-10, atol=1e-13)
    plate = [r for r in open("output/histories/plates_0.dat").read().splitlines()
             if not r.startswith("#")]
    assert [int(r.split()[0]) for r in plate] == [0, 2, 4, 6]
    s.close()
