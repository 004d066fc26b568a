"""Worker of tests/test_gpu_multirank.py: one process per GPU (torchrun).

Runs a case of tests/cases.py decomposed in x-slabs across WORLD_SIZE ranks,
then rank 0 stitches the slabs and compares rho, u and pop_new with the CPU
oracle run on the undecomposed domain.  Exit code 0 = parity.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402
from pylabolt_b200 import capi  # noqa: E402
from pylabolt_b200.comm import SingleComm, TorchComm  # noqa: E402
from pylabolt_b200.io_operator import strip_ghost  # noqa: E402
from pylabolt_b200.operators import (CollisionOperator, FluidLB,  # noqa: E402
                                     ForceOperator)
from pylabolt_b200.solver import Solver  # noqa: E402
from pylabolt_b200.state import State  # noqa: E402

CASES = {
    "cavity": lambda: cases.cavity(37, 29),
    "poiseuille": lambda: cases.poiseuille(26, 21),
    "cylinder_cut": lambda: _centered(cases.cylinder(64, 31, radius=5)),
    "periodic_box": lambda: cases.periodic_box(30, 22),
    "mrt_box": lambda: _mrt(cases.periodic_box(30, 22)),
    "spin": lambda: _centered(cases.cylinder(64, 31, radius=5, spin=0.01)),
    # slabs of 2 (four ranks) / 4 (two ranks) columns: every column is an edge
    "thin": lambda: cases.poiseuille(8, 21),
    # uneven ceil split (parallel/domain.py:54-66): 19 + 18, or 10 + 10 + 10 + 7
    "uneven": lambda: cases.cavity(37, 29),
    # wide enough for fully deep warp strips in every slab (fused path)
    "wide_channel": lambda: _mrt(cases.poiseuille(48, 140)),
    "wide_cylinder": lambda: _centered(cases.cylinder(64, 131, radius=9)),
}


def _centered(sim):
    """Put the body on the slab cut (x = 32 of 64, two ranks)."""
    sim.obstacle_dict["cyl"]["center"] = [32, sim.mesh_dict["grid"][1] // 2]
    return sim


def _mrt(sim):
    sim.collision_dict["fluid"]["model"] = "MRT"
    return sim


def global_oracle(sim):
    from oracle.oracle import Oracle
    comm = SingleComm()
    sim.decompose_dict = {"nx": 1, "ny": 1}
    st = State(sim, comm, 0, verbose=False)
    col = CollisionOperator(sim, FluidLB(), st, comm, verbose=False)
    frc = ForceOperator(sim, FluidLB(), st, comm, collision_operator=col,
                        verbose=False)
    elements = [{"type": el.type_fluid, "nodes": el.boundary_nodes,
                 "out": el.out_list, "inv": el.inv_list, "normal": el.normal,
                 "vector": el.vector_fluid, "scalar": float(el.scalar_fluid)}
                for el in st.boundary.boundary_elements]
    orc = Oracle(st.domain.shape, st.fields.solid, st.fields.ghost_node,
                 st.fields.density, st.fields.velocity, elements,
                 col.omega_fluid, gravity=frc.gravity, forcing=col.forcing_fluid,
                 collision=col.collision_fluid,
                 x_periodic=st.boundary.x_periodic,
                 y_periodic=st.boundary.y_periodic, mrt_rates=col.mrt_rates)
    orc.initialize_pop()
    return orc, st


def fuse_options(tokens):
    """Spec suffixes -> (PLB_FUSE, PLB_FUSE_DEPTH): "fuse2" / "fuse3" / "fuse4"
    force two / three / four steps per pass on any lattice, "fuse0" single
    steps only."""
    if "fuse4" in tokens:
        return "2", "4"
    if "fuse3" in tokens:
        return "2", "3"
    if "fuse2" in tokens:
        return "2", "2"
    if "fuse0" in tokens:
        return "0", "2"
    return "1", "2"


def run_one(comm, name, steps, strict, face, out_dir, fuse="1", depth="2"):
    """One decomposed run against the global oracle; returns 0 on parity."""
    rank, world = comm.Get_rank(), comm.Get_size()
    os.environ["PLB_FACE"] = face
    os.environ["PLB_FUSE"] = fuse
    os.environ["PLB_FUSE_DEPTH"] = depth
    sim = CASES[name]()
    sim.decompose_dict = {"nx": world, "ny": 1}
    solver = Solver(comm, "b200", simulation=sim, strict=strict, verbose=False,
                    device=int(os.environ.get("LOCAL_RANK", "0")))
    solver.set_backend()
    solver.compile()
    transport = solver.plb.info()["faces"]
    solver.plb.initialize_pop()
    solver.advance(steps, store_moments_last=True)
    solver.plb.sync()
    finfo = solver.plb.fused_info()
    pairs = finfo["pairs"] + finfo["triples"] + finfo["quads"]
    if depth == "3" and finfo["n_deep3"] > 0 and steps > 3 and finfo["triples"] == 0:
        pairs = 0
    if depth == "4" and finfo["n_deep4"] > 0 and steps > 4 and finfo["quads"] == 0:
        pairs = 0
    got = solver.fields_to_host()
    shape = solver.state.domain.shape
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"),
             offset=solver.state.domain.offset,
             density=strip_ghost(got["density"], shape),
             velocity=strip_ghost(got["velocity"], shape),
             pop=strip_ghost(got["pop_fluid_new"], shape),
             inner=solver.state.domain.inner_shape)
    res_local = solver.plb.residue_sums()
    res = np.zeros_like(res_local)
    comm.Allreduce(res_local, res)
    solver.close()
    comm.Barrier()
    status = 0
    if rank == 0:
        tag = f"{name} x{world} {'strict' if strict else 'production'} {face}"
        want_transport = {"nccl": capi.FACES_NCCL, "p2p": capi.FACES_P2P}[face]
        if fuse == "2" and pairs == 0:
            print(f"[multirank] {tag}: the fused path did not run", flush=True)
            status = 1
        if transport != want_transport:
            print(f"[multirank] {tag}: face transport {transport}, wanted "
                  f"{want_transport}", flush=True)
            status = 1
        orc, st = global_oracle(CASES[name]())
        orc.step(steps)
        want = {"density": strip_ghost(orc.density, st.domain.shape),
                "velocity": strip_ghost(orc.velocity, st.domain.shape),
                "pop": strip_ghost(orc.pop_new, st.domain.shape)}
        nx, ny = (int(v) for v in st.domain.inner_shape)
        for key, ncomp in (("density", 1), ("velocity", 2), ("pop", 9)):
            full = np.zeros((nx, ny, ncomp))
            for r in range(world):
                d = np.load(os.path.join(out_dir, f"rank{r}.npz"))
                ox = int(d["offset"][0])
                nxr, nyr = (int(v) for v in d["inner"])
                full[ox:ox + nxr] = d[key].reshape(nxr, nyr, ncomp)
            ref = want[key].reshape(nx, ny, ncomp)
            err = np.abs(full - ref).max() / np.abs(ref).max()
            exact = np.array_equal(full, ref)
            print(f"[multirank] {tag} {key}: rel err {err:.3e} "
                  f"bit-exact={exact}", flush=True)
            bgk = sim.collision_dict["fluid"]["model"] == "BGK"
            if not err <= 1e-12 or (strict and bgk and not exact):
                status = 1
        want_res = orc.residue_sums()
        if not np.allclose(res, want_res, rtol=1e-10, atol=1e-300):
            print("[multirank] residue sums differ", res, want_res, flush=True)
            status = 1
    flag = np.array([status], dtype=np.int64)
    out = np.zeros_like(flag)
    comm.Allreduce(flag, out, op="max")
    comm.Barrier()
    return int(out[0])


def main():
    """argv: <out_dir> <name:steps:mode:face> [...] -- all cases in ONE
    torchrun launch (process start-up dominates a case); rank 0 writes
    results.json = {spec: 0 | 1}."""
    import json
    out_dir, specs = sys.argv[1], sys.argv[2:]
    comm = TorchComm()
    results = {}
    for spec in specs:
        name, steps, mode, face, *rest = spec.split(":")
        try:
            fuse, depth = fuse_options(rest)
            results[spec] = run_one(comm, name, int(steps), mode == "strict",
                                    face, out_dir, fuse=fuse, depth=depth)
        except Exception as e:           # noqa: BLE001 -- report, then stop:
            print(f"[multirank] {spec}: {type(e).__name__}: {e}", flush=True)
            results[spec] = 2            # the ranks are no longer in step
            break
    if comm.Get_rank() == 0:
        with open(os.path.join(out_dir, "results.json"), "w") as f:
            json.dump(results, f)
    import torch.distributed as dist
    dist.destroy_process_group()
    sys.exit(max(results.values()) if results else 1)


if __name__ == "__main__":
    main()
