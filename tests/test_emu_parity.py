"""Kernel-logic tests without a GPU (test infrastructure, see tests/emu/README.md).

The unchanged CUDA sources of libplb are compiled by g++ against a stand-in
CUDA runtime whose kernel launches run the threads as cooperative fibers
(warp shuffles, votes, __syncthreads), and the result is driven through the
same C ABI and the same Python host as the GPU.  -ffp-contract=off gives the
arithmetic of the -fmad=false build, so BGK paths must equal the reference's
golden vectors bit for bit.  What is checked here is what needs no hardware:
indexing, the shuffle-assembled 128-bit store patterns, warp-edge cases, the
node classification, the list passes and the host-side ordering of the
single-step path and of the two-steps-per-pass path (k_bulk_fused /
step_fused).  The `-m gpu` tests remain the parity tests proper.
"""
import os
import sys

import numpy as np
import pytest

import cases
from oracle.oracle import Oracle
from pylabolt_b200 import capi
from pylabolt_b200.comm import SingleComm
from pylabolt_b200.solver import Solver

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402

RTOL = 1e-12


@pytest.fixture(scope="module")
def emu_lib():
    return build_emu.build()


@pytest.fixture
def emu(emu_lib, monkeypatch):
    """Routes THIS TEST's solvers to the emulated library (PLB_LIB is the
    loader's explicit override; the product default never points here)."""
    monkeypatch.setenv("PLB_LIB", emu_lib)
    monkeypatch.setattr(capi, "_accept_emulated_build", True)
    monkeypatch.delenv("PLB_FUSED_ROWS", raising=False)
    # the pair-counting tests below pin two steps per pass; the shipped
    # default (three) has its own tests and test_default_is_three_steps_per_pass
    monkeypatch.setenv("PLB_FUSE_DEPTH", "2")
    monkeypatch.delenv("PLB_EMU_BLOCK_ORDER", raising=False)
    monkeypatch.delenv("PLB_FUSED_DYNAMIC", raising=False)
    monkeypatch.delenv("PLB_KERNEL", raising=False)
    return monkeypatch


def rel_err(a, b):
    scale = np.abs(b).max()
    return 0.0 if scale == 0 else float(np.abs(a - b).max() / scale)


def make_solver(sim):
    s = Solver(SingleComm(), "b200", simulation=sim, strict=True, verbose=False)
    s.set_backend()
    s.compile()
    s.plb.initialize_pop()
    return s


def oracle_for(solver, n_threads=4):
    st = solver.state
    col = solver.collision_operator
    elements = [{"type": el.type_fluid, "nodes": el.boundary_nodes,
                 "out": el.out_list, "inv": el.inv_list, "normal": el.normal,
                 "vector": el.vector_fluid, "scalar": float(el.scalar_fluid)}
                for el in st.boundary.boundary_elements]
    orc = Oracle(st.domain.shape, st.fields.solid, st.fields.ghost_node,
                 st.fields.density, st.fields.velocity, elements,
                 col.omega_fluid, gravity=solver.force_operator.gravity,
                 forcing=col.forcing_fluid, collision=col.collision_fluid,
                 x_periodic=st.boundary.x_periodic,
                 y_periodic=st.boundary.y_periodic, mrt_rates=col.mrt_rates,
                 n_threads=n_threads)
    orc.initialize_pop()
    return orc


# PLB_FUSE: 0 = single steps only, 2 = two steps per pass wherever a deep node
# exists (the default, 1, skips lattices as small as the golden ones)
@pytest.mark.parametrize("fuse", ["0", "2"])
@pytest.mark.parametrize("variant", ["vec2", "scalar"])
@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CASES))
def test_emulated_kernels_are_bit_exact_with_reference(golden_dir, emu, name,
                                                       variant, fuse):
    if fuse == "2" and variant == "scalar":
        pytest.skip("the fused path has one kernel variant")
    emu.setenv("PLB_KERNEL", variant)
    emu.setenv("PLB_FUSE", fuse)
    factory, kwargs, record = cases.GOLDEN_CASES[name]
    data = np.load(os.path.join(golden_dir, name + ".npz"))
    s = make_solver(factory(**kwargs))
    try:
        assert np.array_equal(s.plb.download(capi.POP), data["pop_0"])
        done = 0
        for step in record:
            s.advance(step - done, store_moments_last=True)
            done = step
            got = s.fields_to_host()
            assert np.array_equal(got["density"], data[f"density_{step}"]), step
            assert np.array_equal(got["velocity"], data[f"velocity_{step}"]), step
            assert np.array_equal(got["pop_fluid_new"], data[f"pop_{step}"]), step
        info = s.plb.fused_info()
        assert (info["pairs"] > 0) == (fuse == "2")
    finally:
        s.close()


def _mrt(sim):
    sim.collision_dict["fluid"]["model"] = "MRT"
    return sim


def _mrt_free_rates(sim):
    sim.collision_dict["fluid"]["model"] = "MRT"
    sim.collision_dict["fluid"]["mrt_rates"] = [1.0, 1.4, 1.3, 1.0, 1.2, 1.0,
                                                1.2, 1.7, 1.6]
    return sim


def _zero_gradient_outlet(sim):
    sim.boundary_dict["outlet"]["fluid"] = {"type": "zero_gradient"}
    return sim


# Lattices wide enough that whole 62-node strips are deep (the 128-bit store
# path of the fused kernel), several strips, several row chunks, odd sizes,
# the periodic seam, bodies, every collision / forcing model.
WIDE_CASES = {
    "poiseuille_70x140_guo2": lambda: cases.poiseuille(70, 140),
    "poiseuille_33x190_guo1": lambda: cases.poiseuille(33, 190, forcing="guo_linear"),
    "poiseuille_9x300_none": lambda: cases.poiseuille(9, 300, forcing=None),
    "mrt_poiseuille_70x140_guo2": lambda: _mrt(cases.poiseuille(70, 140)),
    "mrt_poiseuille_40x129_guo1": lambda: _mrt(cases.poiseuille(40, 129, forcing="guo_linear")),
    "mrt_free_rates_45x131": lambda: _mrt_free_rates(cases.periodic_box(45, 131)),
    "periodic_box_67x131": lambda: cases.periodic_box(67, 131),
    "cavity_40x200": lambda: cases.cavity(40, 200),
    "cylinder_120x140": lambda: cases.cylinder(120, 140),
    "channel_y_140x40": lambda: cases.channel_y(140, 40),
    "zero_gradient_100x127": lambda: _zero_gradient_outlet(cases.inflow_cylinder(100, 127)),
    "mrt_zero_gradient_64x125": lambda: _mrt(_zero_gradient_outlet(cases.inflow_cylinder(64, 125))),
    # degenerate: too thin for any deep node -> the pair path must stay off
    "three_columns_periodic": lambda: cases.poiseuille(3, 140),
    "two_rows_y_periodic": lambda: cases.channel_y(140, 2),
}


def _run(sim_factory, n_steps, fuse, emu, rows=None, one_by_one=False, depth=2):
    emu.setenv("PLB_FUSE", fuse)
    emu.setenv("PLB_FUSE_DEPTH", str(depth))
    if rows is None:
        emu.delenv("PLB_FUSED_ROWS", raising=False)
    else:
        emu.setenv("PLB_FUSED_ROWS", str(rows))
    # CTAs last to first for the chunked runs: a result that depended on the
    # order of the CTAs would have two writers for one slot
    emu.setenv("PLB_EMU_BLOCK_ORDER", "reverse" if rows in (5, 7) else "forward")
    # ... and the one-step-per-call runs draw their work items from the queue
    # of the persistent-grid variant (PLB_FUSED_DYNAMIC=1)
    emu.setenv("PLB_FUSED_DYNAMIC", "1" if one_by_one else "0")
    s = make_solver(sim_factory())
    try:
        if one_by_one:
            for _ in range(n_steps - 1):
                s.execute_single_time_step()
            s.single_time_step(store_moments=True)
        else:
            s.advance(n_steps, store_moments_last=True)
        return s.fields_to_host(), s.plb.fused_info()
    finally:
        s.close()


@pytest.mark.parametrize("name", sorted(WIDE_CASES))
def test_two_steps_per_pass_equals_two_single_steps(emu, name):
    """Same per-node arithmetic, so the fused path must reproduce the
    single-step path BIT FOR BIT (any collision model), for even and odd step
    counts, for several chunk heights, and when the host issues one step per
    call like the reference's Solver.run."""
    factory = WIDE_CASES[name]
    want, info0 = _run(factory, 13, "0", emu)
    assert info0["pairs"] == 0
    for rows, one_by_one in ((None, False), (5, False), (64, True)):
        got, info = _run(factory, 13, "2", emu, rows=rows, one_by_one=one_by_one)
        if info["n_deep"] == 0:
            assert info["active"] == 0 and info["pairs"] == 0
        else:
            assert info["pairs"] == 6, info
        for key in ("density", "velocity", "pop_fluid_new"):
            assert np.array_equal(got[key], want[key]), (name, rows, key)


@pytest.mark.parametrize("name", sorted(WIDE_CASES))
def test_three_steps_per_pass_equals_three_single_steps(emu, name):
    """PLB_FUSE_DEPTH=3: the same kernel template one level deeper (60 of 64
    nodes per warp strip, chunks overlap by four rows, three list passes, two
    scratch lattices); 14 steps = 4 triples + 1 pair."""
    factory = WIDE_CASES[name]
    want, _ = _run(factory, 15, "0", emu)
    for rows, one_by_one in ((None, False), (7, False), (64, True)):
        got, info = _run(factory, 15, "2", emu, rows=rows, one_by_one=one_by_one,
                         depth=3)
        if info["n_deep3"] > 0:
            assert info["active"] == 3 and info["triples"] == 4 and info["pairs"] == 1, info
        elif info["n_deep"] > 0:
            assert info["active"] == 2 and info["pairs"] == 7, info
        else:
            assert info["active"] == 0
        for key in ("density", "velocity", "pop_fluid_new"):
            assert np.array_equal(got[key], want[key]), (name, rows, key)


@pytest.mark.parametrize("name", sorted(WIDE_CASES))
def test_four_steps_per_pass_equals_four_single_steps(emu, name):
    """PLB_FUSE_DEPTH=4: one more level (58 of 64 nodes per warp strip, chunks
    overlap by six rows, four list passes, three scratch lattices, deep flags
    up to 3); 14 plain steps = 3 groups of four + 1 pair."""
    factory = WIDE_CASES[name]
    want, _ = _run(factory, 15, "0", emu)
    for rows, one_by_one in ((None, False), (7, False), (64, True)):
        got, info = _run(factory, 15, "2", emu, rows=rows, one_by_one=one_by_one,
                         depth=4)
        if info["n_deep4"] > 0:
            assert info["active"] == 4 and info["quads"] == 3 and info["pairs"] == 1, info
            assert info["n_deep4"] < info["n_deep3"] < info["n_deep"]
        elif info["n_deep3"] > 0:
            assert info["active"] == 3 and info["triples"] == 4 and info["pairs"] == 1, info
        elif info["n_deep"] > 0:
            assert info["active"] == 2 and info["pairs"] == 7, info
        else:
            assert info["active"] == 0
        for key in ("density", "velocity", "pop_fluid_new"):
            assert np.array_equal(got[key], want[key]), (name, rows, key)


@pytest.mark.parametrize("depth", [3, 4])
@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CASES))
def test_three_steps_per_pass_is_bit_exact_with_reference(golden_dir, emu, name, depth):
    """The reference's own runs (golden vectors), three and four steps per pass."""
    emu.setenv("PLB_FUSE", "2")
    emu.setenv("PLB_FUSE_DEPTH", str(depth))
    factory, kwargs, record = cases.GOLDEN_CASES[name]
    data = np.load(os.path.join(golden_dir, name + ".npz"))
    s = make_solver(factory(**kwargs))
    try:
        done = 0
        for step in record:
            s.advance(step - done, store_moments_last=True)
            done = step
            got = s.fields_to_host()
            assert np.array_equal(got["density"], data[f"density_{step}"]), step
            assert np.array_equal(got["velocity"], data[f"velocity_{step}"]), step
            assert np.array_equal(got["pop_fluid_new"], data[f"pop_{step}"]), step
        info = s.plb.fused_info()
        assert info["triples"] + info["quads"] > 0
        assert info["active"] <= depth
    finally:
        s.close()


@pytest.mark.parametrize("name", ["poiseuille_70x140_guo2", "cylinder_120x140",
                                  "mrt_poiseuille_70x140_guo2",
                                  "zero_gradient_100x127"])
def test_fused_path_against_oracle(emu, name):
    emu.setenv("PLB_FUSE", "2")
    s = make_solver(WIDE_CASES[name]())
    try:
        orc = oracle_for(s)
        for n in (2, 7, 16):
            s.advance(n, store_moments_last=True)
            orc.step(n)
            got = s.fields_to_host()
            assert rel_err(got["density"], orc.density) <= RTOL
            assert rel_err(got["velocity"], orc.velocity) <= RTOL
            assert rel_err(got["pop_fluid_new"], orc.pop_new) <= RTOL
            bgk = s.collision_operator.collision_fluid == "BGK"
            if bgk and "zero_gradient" not in name:
                assert np.array_equal(got["pop_fluid_new"], orc.pop_new)
        assert s.plb.fused_info()["pairs"] == (0 + 3 + 7)   # flagged last steps run unfused
    finally:
        s.close()


def test_default_mode_pairs_only_where_deep_nodes_dominate(emu):
    emu.delenv("PLB_FUSE", raising=False)
    small = make_solver(cases.cavity())              # 33 x 29: list passes dominate
    large = make_solver(cases.cavity(101, 101))      # BASELINE configs[0]
    try:
        assert small.plb.fused_info()["active"] == 0
        info = large.plb.fused_info()
        assert info["active"] == 2 and info["n_deep"] == 97 * 97
        large.advance(10)
        large.plb.sync()
        assert large.plb.fused_info()["pairs"] == 5
    finally:
        small.close()
        large.close()


@pytest.mark.parametrize("depth", [2, 3, 4])
def test_chunks_marching_both_ways_equal_chunks_marching_one_way(emu, depth):
    """Odd chunks march towards smaller x by default (their neighbours then
    reach the shared halo rows at the same time); PLB_FUSED_ALTERNATE=0 sends
    every chunk towards larger x.  Same fields, bit for bit, and equal to
    single steps -- chunks of 3 rows: shorter than the halo at depth 4."""
    factory = WIDE_CASES["mrt_poiseuille_70x140_guo2"]
    want, _ = _run(factory, 13, "0", emu)
    # bit 0: odd chunks march towards smaller x; bit 1: the warps of a CTA take
    # x-adjacent chunks of one strip (also with the work queue: items past the
    # last chunk of a group are followed by valid ones)
    for alternate in ("1", "0", "3", "2"):
        emu.setenv("PLB_FUSED_ALTERNATE", alternate)
        for rows in (3, 8):
            got, info = _run(factory, 13, "2", emu, rows=rows, depth=depth,
                             one_by_one=(alternate == "3" and rows == 8))
            assert info["pairs"] + info["triples"] + info["quads"] > 0
            for key in ("density", "velocity", "pop_fluid_new"):
                assert np.array_equal(got[key], want[key]), (alternate, rows, key)
    emu.delenv("PLB_FUSED_ALTERNATE", raising=False)


def test_default_is_three_steps_per_pass(emu):
    """Shipped default for BGK: plain steps go three at a time, a remainder of
    two as a pair, a single one through the single-step kernel."""
    emu.delenv("PLB_FUSE", raising=False)
    emu.delenv("PLB_FUSE_DEPTH", raising=False)
    factory = lambda: cases.cavity(201, 201)
    s = make_solver(factory())
    try:
        info = s.plb.fused_info()
        assert info["active"] == 3 and info["n_deep3"] == 195 * 195
        s.advance(11)
        s.plb.sync()
        info = s.plb.fused_info()
        assert info["triples"] == 3 and info["pairs"] == 1
        s.advance(1, store_moments_last=True)
        got = s.fields_to_host()
    finally:
        s.close()
    want, _ = _run(factory, 12, "0", emu)
    for key in ("density", "velocity", "pop_fluid_new"):
        assert np.array_equal(got[key], want[key]), key


def test_default_is_four_steps_per_pass_for_the_mrt_shortcut(emu):
    """MRT with the reference's rates (the two-stress-moment kernel) groups
    plain steps four at a time by default, a remainder of three as one
    three-step pass."""
    emu.delenv("PLB_FUSE", raising=False)
    emu.delenv("PLB_FUSE_DEPTH", raising=False)
    factory = lambda: _mrt(cases.poiseuille(210, 200))
    s = make_solver(factory())
    try:
        info = s.plb.fused_info()
        assert info["active"] == 4 and 0 < info["n_deep4"] < info["n_deep3"], info
        s.advance(11)
        s.plb.sync()
        info = s.plb.fused_info()
        assert info["quads"] == 2 and info["triples"] == 1 and info["pairs"] == 0, info
        s.advance(1, store_moments_last=True)
        got = s.fields_to_host()
    finally:
        s.close()
    want, _ = _run(factory, 12, "0", emu)
    for key in ("density", "velocity", "pop_fluid_new"):
        assert np.array_equal(got[key], want[key]), key


def test_held_back_step_is_completed_by_every_observer(emu):
    """plb_step may hold one plain step back for pairing; download, sync,
    residues and a flagged step must all see it done."""
    emu.setenv("PLB_FUSE", "2")
    factory = WIDE_CASES["poiseuille_70x140_guo2"]
    want, _ = _run(factory, 5, "0", emu)
    emu.setenv("PLB_FUSE", "2")
    s = make_solver(factory())
    try:
        for _ in range(3):
            s.plb.step(1)                     # 2 run as a pair, 1 is held back
        assert s.plb.fused_info()["pairs"] == 1
        pop3 = s.plb.download(capi.POP)       # completes the third step
        s.plb.step(1)
        s.plb.step(1, store_moments=True)     # flagged: runs 4 and 5 unfused
        got = s.fields_to_host()
        for key in ("density", "velocity", "pop_fluid_new"):
            assert np.array_equal(got[key], want[key]), key
        ref3, _ = _run(factory, 3, "0", emu)
        assert np.array_equal(pop3, ref3["pop_fluid_new"])
    finally:
        s.close()


def test_residues_and_link_record_with_pairing(emu):
    emu.setenv("PLB_FUSE", "0")
    a = make_solver(cases.cylinder(120, 140))
    emu.setenv("PLB_FUSE", "2")
    b = make_solver(cases.cylinder(120, 140))
    try:
        for s in (a, b):
            s.advance(6, store_moments_last=True)
            s.plb.residue_sums()
            s.plb.step(5, store_moments=True, record_links=True)
        assert np.array_equal(a.plb.residue_sums(), b.plb.residue_sums())
        n = a.plb.info()["n_link"]
        assert np.array_equal(a.plb.link_exchange(n), b.plb.link_exchange(n))
        assert b.plb.fused_info()["pairs"] == 2 + 2
    finally:
        a.close()
        b.close()


# ---- tuning variants of the fused kernel (tools/build_variants.py) ------------
VARIANT_FLAGS = {
    # PLB_FUSED_BULK=1: the prefetch ring filled by TMA bulk copies that
    # complete on per-warp mbarriers (three slots)
    "bulk_s3": ["-DPLB_FUSED_CARRY_SMEM=0", "-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=3",
                "-DPLB_FUSED_TENSOR=0"],
    # ... and PLB_FUSED_CARRY_SMEM=1: the carried populations in shared memory
    # (everything in dynamic shared memory), a ring of one slot, refilled as
    # soon as it has been read (round 2's first shipped configuration)
    "carry_bulk_s1": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1",
                      "-DPLB_FUSED_STAGES=1", "-DPLB_FUSED_TENSOR=0"],
    # PLB_FUSED_TENSOR=1 (the shipped default with one slot): a warp's row is one
    # rank-3 tensor copy into a dense per-warp slot; here with three slots
    "tensor_s3": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1",
                  "-DPLB_FUSED_STAGES=3", "-DPLB_FUSED_TENSOR=1"],
}


@pytest.fixture(scope="module", params=sorted(VARIANT_FLAGS))
def emu_variant_lib(request):
    return build_emu.build(variant=request.param,
                           extra_flags=VARIANT_FLAGS[request.param])


@pytest.mark.parametrize("name", ["mrt_poiseuille_70x140_guo2", "cylinder_120x140",
                                  "poiseuille_9x300_none"])
def test_fused_kernel_variants_equal_single_steps(emu, emu_variant_lib, name):
    """Slot / phase bookkeeping of the mbarrier ring and the odd / even row
    slots of the shared-memory carry: chunks shorter than the ring, work items
    drawn from the queue (the running fill count crosses items, stale carry
    slots of the previous item), both depths -- bit for bit against the
    single-step path."""
    factory = WIDE_CASES[name]
    want, _ = _run(factory, 15, "0", emu)
    emu.setenv("PLB_LIB", emu_variant_lib)
    for rows, one_by_one, depth in ((None, False, 2), (1, False, 2), (64, True, 2),
                                    (7, True, 3)):
        got, info = _run(factory, 15, "2", emu, rows=rows, one_by_one=one_by_one,
                         depth=depth)
        assert info["pairs"] + info["triples"] > 0, info
        for key in ("density", "velocity", "pop_fluid_new"):
            assert np.array_equal(got[key], want[key]), (name, rows, depth, key)


def test_uniform_initial_fields_are_filled_on_the_device(emu):
    """A field the case file fixes to one value is produced by plb_fill; the
    device then holds exactly what an upload of the host array would."""
    sim = cases.cavity(40, 33)
    sim.initial_fields_dict["default"]["fluid"]["velocity"]["value"] = [0.02, -0.01]
    s = make_solver(sim)
    try:
        assert s._uniform_initial_value("density") == 1.0
        assert s._uniform_initial_value("velocity") == [0.02, -0.01]
        assert s.upload_initial_fields() == 0              # nothing crossed PCIe
        assert np.array_equal(s.plb.download(capi.DENSITY), s.state.fields.density)
        assert np.array_equal(s.plb.download(capi.VELOCITY), s.state.fields.velocity)
    finally:
        s.close()
    body = make_solver(cases.cylinder())                  # a body: uploads
    try:
        assert body._uniform_initial_value("density") is None
        assert body.upload_initial_fields() == (body.state.fields.density.nbytes +
                                                body.state.fields.velocity.nbytes)
    finally:
        body.close()


def random_bodies_case(seed, nx=150, ny=200):
    """Channel with pressure in / outlet and five circles / inclined ellipses
    thrown in at random (non-overlapping, as the reference demands), some of
    them closer to each other and to the walls than the radius a four-step
    pass needs."""
    rng = np.random.default_rng(seed)
    sim = cases.cylinder(nx, ny, radius=6)
    bodies = {"options": {}}
    placed = []
    for b in range(5):
        for _ in range(50):
            r = int(rng.integers(3, 9))
            cx = int(rng.integers(r + 3, nx - r - 3))
            cy = int(rng.integers(r + 3, ny - r - 3))
            if all((cx - px) ** 2 + (cy - py) ** 2 > (r + pr + 2) ** 2
                   for px, py, pr in placed):
                placed.append((cx, cy, r))
                break
        else:
            continue
        if b % 2:
            bodies[f"body{b}"] = {"type": "circle", "radius": r, "center": [cx, cy],
                                  "density": 1.0, "static": True}
        else:
            bodies[f"body{b}"] = {"type": "ellipse", "semi_major_axis": float(r),
                                  "semi_minor_axis": float(max(2, r - 2)),
                                  "inclination_angle": float(rng.integers(0, 90)),
                                  "center": [cx, cy], "density": 1.0, "static": True}
    sim.obstacle_dict = bodies
    return sim


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_bodies_every_depth_equals_single_steps(emu, seed):
    """The deep flags, the list passes of every depth and the fused kernel
    around randomly placed bodies must add up to single steps, bit for bit."""
    factory = lambda: random_bodies_case(seed)
    want, _ = _run(factory, 10, "0", emu)
    for depth in (2, 3, 4):
        got, info = _run(factory, 10, "2", emu, rows=9, depth=depth)
        assert info["pairs"] + info["triples"] + info["quads"] > 0, info
        for key in ("density", "velocity", "pop_fluid_new"):
            assert np.array_equal(got[key], want[key]), (seed, depth, key)
