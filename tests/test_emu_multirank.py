"""Multi-rank logic of libplb without a GPU (test infrastructure, see
tests/emu/README.md): every rank is a host thread driving its own solver
through the emulated library; slab faces travel through the library's real
code paths -- peer-to-peer stores into the neighbour's receive buffer with
the mailbox hand-shake ("IPC" handles are pointers here), or the NCCL
send / receive sequence served by an in-process stand-in -- and an operation
whose neighbour has not arrived yet blocks, as on the device.  The stitched
result must be BIT-IDENTICAL to the CPU oracle on the undecomposed domain
(the same check, by the same function, as tests/test_gpu_multirank.py).

Covers what a single GPU box cannot: both face transports, 2 / 3 / 4 slabs,
bodies cut by a slab face, the periodic seam through the ring of ranks, uneven
splits, and the two-steps-per-pass path (":fuse2") whose list passes exchange
faces twice per pass.
"""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
import build_emu  # noqa: E402

SPECS = {
    2: ["cavity:25:strict:p2p", "cylinder_cut:25:strict:nccl",
        "thin:60:strict:p2p",
        "periodic_box:25:strict:p2p:fuse2", "cylinder_cut:26:strict:nccl:fuse2",
        "spin:25:strict:p2p:fuse2", "mrt_box:25:strict:p2p:fuse2",
        "poiseuille:101:strict:nccl:fuse2", "wide_channel:41:strict:p2p:fuse2",
        "wide_cylinder:40:strict:p2p:fuse2",
        # three steps per pass: three list passes, three exchanges per pass
        "periodic_box:25:strict:p2p:fuse3", "cylinder_cut:26:strict:nccl:fuse3",
        "wide_channel:41:strict:p2p:fuse3", "wide_cylinder:40:strict:nccl:fuse3",
        "wide_channel:42:strict:p2p:fuse4", "wide_cylinder:41:strict:nccl:fuse4",
        "periodic_box:25:strict:p2p:fuse4"],
    3: ["wide_cylinder:40:strict:p2p:fuse2", "poiseuille:25:strict:nccl:fuse2",
        "wide_channel:40:strict:p2p:fuse3"],
    4: ["uneven:41:strict:p2p:fuse2", "periodic_box:25:strict:p2p:fuse2",
        "wide_channel:60:strict:nccl:fuse2", "thin:60:strict:nccl",
        "uneven:41:strict:p2p:fuse3", "wide_cylinder:61:strict:p2p:fuse3",
        "wide_channel:61:strict:p2p:fuse4"],
}


def _launch(world, tmp_path_factory):
    lib = build_emu.build()
    out = str(tmp_path_factory.mktemp(f"emu_world{world}") / "results.json")
    env = dict(os.environ, PLB_LIB=lib, PLB_P2P_TIMEOUT_S="20",
               LD_LIBRARY_PATH=os.path.dirname(lib) + os.pathsep +
               os.environ.get("LD_LIBRARY_PATH", ""))
    env.pop("PLB_FUSED_ROWS", None)
    proc = subprocess.run(
        [sys.executable, os.path.join(HERE, "emu", "multirank_emu_worker.py"),
         str(world), out] + SPECS[world],
        capture_output=True, text=True, timeout=1500, env=env)
    log = proc.stdout[-6000:] + "\n" + proc.stderr[-3000:]
    try:
        with open(out) as f:
            return json.load(f), log
    except OSError:
        return {"results": {}, "failures": ["no result file"]}, log


@pytest.fixture(scope="module")
def world2(tmp_path_factory):
    return _launch(2, tmp_path_factory)


@pytest.fixture(scope="module")
def world3(tmp_path_factory):
    return _launch(3, tmp_path_factory)


@pytest.fixture(scope="module")
def world4(tmp_path_factory):
    return _launch(4, tmp_path_factory)


@pytest.mark.parametrize("spec", SPECS[2])
def test_two_emulated_slabs_match_oracle(world2, spec):
    data, log = world2
    assert not data["failures"], log
    assert data["results"].get(spec) == 0, log


@pytest.mark.parametrize("spec", SPECS[3])
def test_three_emulated_slabs_match_oracle(world3, spec):
    data, log = world3
    assert not data["failures"], log
    assert data["results"].get(spec) == 0, log


@pytest.mark.parametrize("spec", SPECS[4])
def test_four_emulated_slabs_match_oracle(world4, spec):
    data, log = world4
    assert not data["failures"], log
    assert data["results"].get(spec) == 0, log
