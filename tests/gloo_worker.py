"""CPU worker of tests/test_gloo_multirank.py (torchrun, gloo backend).

Each rank builds ITS slab with pylabolt_b200's host layer (State, boundary
link lists, obstacle flags incl. ghost columns) and advances it with the CPU
oracle's individual phases; between collision and streaming the ranks
exchange the post-collision ghost columns over gloo -- the reference's
halo_exchange (parallel/MPI_operator.py:155-259) for x-slabs.  Rank 0 then
checks that the stitched result is BIT-IDENTICAL to the oracle on the
undecomposed domain: the multi-rank setup (ownership of boundary nodes, solid
flags and wall velocities in ghost columns, neighbour ranks, reductions) is
decomposition independent.
"""
import ctypes
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from pylabolt_b200.comm import SingleComm, TorchComm  # noqa: E402
from pylabolt_b200.io_operator import strip_ghost  # noqa: E402
from pylabolt_b200.operators import (CollisionOperator, FluidLB,  # noqa: E402
                                     ForceOperator)
from pylabolt_b200.solver import neighbour_ranks  # noqa: E402
from pylabolt_b200.state import State  # noqa: E402

CASES = {
    "cavity": lambda: cases.cavity(37, 29),
    "poiseuille": lambda: cases.poiseuille(26, 21),
    "periodic_box": lambda: cases.periodic_box(30, 22),
    "spin_cut": lambda: _centered(cases.cylinder(64, 31, radius=5, spin=0.02)),
}


def _centered(sim):
    sim.obstacle_dict["cyl"]["center"] = [32, 15]
    return sim


def build_oracle(sim, comm, rank):
    st = State(sim, comm, rank, verbose=False)
    col = CollisionOperator(sim, FluidLB(), st, comm, verbose=False)
    frc = ForceOperator(sim, FluidLB(), st, comm, collision_operator=col,
                        verbose=False)
    st.obstacle.check_overlap(comm)
    elements = [{"type": el.type_fluid, "nodes": el.boundary_nodes,
                 "out": el.out_list, "inv": el.inv_list, "normal": el.normal,
                 "vector": el.vector_fluid, "scalar": float(el.scalar_fluid)}
                for el in st.boundary.boundary_elements]
    # x wrap / neighbours are handled by the explicit exchange below
    orc = Oracle(st.domain.shape, st.fields.solid, st.fields.ghost_node,
                 st.fields.density, st.fields.velocity, elements,
                 col.omega_fluid, gravity=frc.gravity, forcing=col.forcing_fluid,
                 collision=col.collision_fluid, x_periodic=False,
                 y_periodic=st.boundary.y_periodic, n_threads=2)
    orc.initialize_pop()
    return orc, st


def exchange_columns(orc, st, left, right):
    """Ghost column x=0 <- left neighbour's last inner column, ghost column
    x=Nx_pad-1 <- right neighbour's first inner column (full columns)."""
    nxp, nyp = orc.nx_pad, orc.ny_pad
    pop = orc.pop.reshape(nxp, nyp, 9)
    ops, bufs = [], {}
    if right is not None:
        send = torch.from_numpy(pop[nxp - 2].copy())
        bufs["right"] = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, right, tag=1),
                dist.P2POp(dist.irecv, bufs["right"], right, tag=2)]
    if left is not None:
        send = torch.from_numpy(pop[1].copy())
        bufs["left"] = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, left, tag=2),
                dist.P2POp(dist.irecv, bufs["left"], left, tag=1)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if "left" in bufs:
        pop[0] = bufs["left"].numpy()
    if "right" in bufs:
        pop[nxp - 1] = bufs["right"].numpy()


def step(orc, st, left, right):
    lib, p, f = orc.lib, ctypes.byref(orc.params), ctypes.byref(orc.fields)
    lib.oracle_density(p, f)
    lib.oracle_gravity_force(p, f)
    lib.oracle_velocity(p, f)
    lib.oracle_collide_bgk(p, f)
    exchange_columns(orc, st, left, right)
    # y wrap on full rows after the x phase, like the reference
    lib.oracle_wrap_ghosts(p, orc.pop.ctypes.data_as(ctypes.c_void_p),
                           ctypes.c_int64(72))
    lib.oracle_stream(p, f)
    lib.oracle_set_boundary(p, f, orc.elements, ctypes.c_int32(orc.n_elements))


def main():
    name, steps, out_dir = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    comm = TorchComm(backend="gloo")
    rank, world = comm.Get_rank(), comm.Get_size()
    assert comm.backend == "gloo"
    # plumbing used by the solver
    token = comm.bcast_bytes(bytes(range(128)) if rank == 0 else bytes(128))
    assert token == bytes(range(128))
    total = np.zeros(3)
    comm.Allreduce(np.array([1.0, rank, 2.5]), total)
    assert total.tolist() == [world, world * (world - 1) / 2, 2.5 * world]
    peak = np.zeros(1)
    comm.Allreduce(np.array([float(rank)]), peak, op="max")
    assert peak[0] == world - 1

    sim = CASES[name]()
    sim.decompose_dict = {"nx": world, "ny": 1}
    orc, st = build_oracle(sim, comm, rank)
    left, right = neighbour_ranks(st.domain, st.boundary)
    if world == 1:
        left = right = None
    for _ in range(steps):
        step(orc, st, left, right)
    res = np.zeros(6)
    comm.Allreduce(orc.residue_sums(), res)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"),
             offset=st.domain.offset, inner=st.domain.inner_shape,
             pop=strip_ghost(orc.pop_new, st.domain.shape),
             density=strip_ghost(orc.density, st.domain.shape))
    comm.Barrier()
    status = 0
    if rank == 0:
        whole_sim = CASES[name]()
        whole, wst = None, None
        single = SingleComm()
        wst = State(whole_sim, single, 0, verbose=False)
        col = CollisionOperator(whole_sim, FluidLB(), wst, single, verbose=False)
        frc = ForceOperator(whole_sim, FluidLB(), wst, single,
                            collision_operator=col, verbose=False)
        elements = [{"type": el.type_fluid, "nodes": el.boundary_nodes,
                     "out": el.out_list, "inv": el.inv_list,
                     "normal": el.normal, "vector": el.vector_fluid,
                     "scalar": float(el.scalar_fluid)}
                    for el in wst.boundary.boundary_elements]
        whole = Oracle(wst.domain.shape, wst.fields.solid,
                       wst.fields.ghost_node, wst.fields.density,
                       wst.fields.velocity, elements, col.omega_fluid,
                       gravity=frc.gravity, forcing=col.forcing_fluid,
                       collision=col.collision_fluid,
                       x_periodic=wst.boundary.x_periodic,
                       y_periodic=wst.boundary.y_periodic)
        whole.initialize_pop()
        whole.step(steps)
        nx, ny = (int(v) for v in wst.domain.inner_shape)
        for key, ncomp, ref in (("pop", 9, whole.pop_new),
                                ("density", 1, whole.density)):
            ref = strip_ghost(ref, wst.domain.shape).reshape(nx, ny, ncomp)
            full = np.zeros_like(ref)
            for r in range(world):
                d = np.load(os.path.join(out_dir, f"rank{r}.npz"))
                ox = int(d["offset"][0])
                nxr, nyr = (int(v) for v in d["inner"])
                full[ox:ox + nxr] = d[key].reshape(nxr, nyr, ncomp)
            if not np.array_equal(full, ref):
                print(f"[gloo] {name}: {key} differs, max "
                      f"{np.abs(full - ref).max():.3e}", flush=True)
                status = 1
        if not np.allclose(res, whole.residue_sums(), rtol=1e-12, atol=1e-300):
            print("[gloo] residue sums differ", flush=True)
            status = 1
        print(f"[gloo] {name} x{world}: status {status}", flush=True)
    flag = np.zeros(1)
    comm.Allreduce(np.array([float(status)]), flag, op="max")
    dist.destroy_process_group()
    sys.exit(int(flag[0]))


if __name__ == "__main__":
    main()
