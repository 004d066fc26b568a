"""Slab decomposition on the host: Domain sizes / offsets, neighbour ranks,
and -- rank by rank with a stand-in communicator, the way the reference tests
multi-rank ownership (tests/unit/factories.py:84-97, test_domain.py:121-187,
test_boundary.py:636-702) -- that every slab's flags, fields and link lists are
exactly the matching slice of the undecomposed state."""
import numpy as np
import pytest

import cases
from pylabolt_b200.solver import neighbour_ranks
from pylabolt_b200.state import Domain, Mesh, State


class DummyComm:
    def __init__(self, rank, size):
        self.rank, self.size = rank, size

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def Barrier(self):
        pass

    def Abort(self, code=1):
        pass

    def Allreduce(self, local, out, op="sum"):
        out[...] = local


@pytest.mark.parametrize("n_global,n_procs", [(100, 1), (100, 3), (101, 4),
                                              (16384, 8), (65536, 8), (17, 9)])
def test_ceil_split_sizes_and_offsets(n_global, n_procs):
    """parallel/domain.py:54-76: ceil(N/n) everywhere but the last rank."""
    from types import SimpleNamespace
    sim = SimpleNamespace(mesh_dict={"grid": [n_global, 8]},
                          decompose_dict={"nx": n_procs, "ny": 1})
    chunk = -(-n_global // n_procs)
    total = 0
    for r in range(n_procs):
        mesh = Mesh(sim, r, verbose=False)
        d = Domain(sim, mesh, DummyComm(r, n_procs))
        assert (d.i_proc, d.j_proc) == (r, 0)
        assert d.offset[0] == r * chunk and d.offset[1] == 0
        want = chunk if r != n_procs - 1 else n_global - r * chunk
        assert d.Nx_rank == want and d.Ny_rank == 8
        assert tuple(d.shape) == (want + 2, 10)
        total += d.Nx_rank
    assert total == n_global


def test_wrong_process_count_is_rejected():
    from types import SimpleNamespace
    sim = SimpleNamespace(mesh_dict={"grid": [16, 8]},
                          decompose_dict={"nx": 3, "ny": 1})
    with pytest.raises(ValueError, match="invalid domain decomposition"):
        Domain(sim, Mesh(sim, 0, verbose=False), DummyComm(0, 2))


def test_y_decomposition_is_refused():
    sim = cases.cavity()
    sim.decompose_dict = {"nx": 1, "ny": 2}
    with pytest.raises(ValueError, match="x-slabs"):
        State(sim, DummyComm(0, 2), 0, verbose=False)


def test_neighbour_ranks():
    """parallel/MPI_operator.py:116-153 for nx = G, ny = 1."""
    from types import SimpleNamespace
    for periodic in (False, True):
        b = SimpleNamespace(x_periodic=periodic)
        for n in (1, 2, 4, 8):
            for r in range(n):
                d = SimpleNamespace(no_of_procs_x=n, i_proc=r)
                left, right = neighbour_ranks(d, b)
                if periodic:
                    assert left == (r - 1) % n and right == (r + 1) % n
                else:
                    assert left == (None if r == 0 else r - 1)
                    assert right == (None if r == n - 1 else r + 1)


SLAB_CASES = {
    "cavity": lambda: cases.cavity(37, 29),
    "poiseuille": lambda: cases.poiseuille(26, 21),
    "cylinder": lambda: cases.cylinder(64, 31, radius=5),
    "periodic_box": lambda: cases.periodic_box(30, 22),
    "ellipse": lambda: cases.inflow_cylinder(48, 27),
    "spin": lambda: cases.cylinder(64, 31, radius=5, spin=0.02),
}


@pytest.mark.parametrize("n_ranks", [2, 3, 5])
@pytest.mark.parametrize("name", sorted(SLAB_CASES))
def test_slabs_are_slices_of_the_global_state(name, n_ranks):
    sim = SLAB_CASES[name]()
    if name in ("cylinder", "spin"):
        sim.obstacle_dict["cyl"]["center"] = [32, 15]
    whole = State(sim, DummyComm(0, 1), 0, verbose=False)
    nxp, nyp = (int(v) for v in whole.domain.shape)
    g = whole.fields

    def grid(a):
        return a.reshape((nxp, nyp) + a.shape[1:])

    sim.decompose_dict = {"nx": n_ranks, "ny": 1}
    owned_nodes = [set() for _ in whole.boundary.boundary_elements]
    for r in range(n_ranks):
        st = State(sim, DummyComm(r, n_ranks), r, verbose=False)
        ox = int(st.domain.offset[0])
        nx = st.domain.Nx_rank
        lx, ly = (int(v) for v in st.domain.shape)
        f = st.fields

        def local(a):
            return a.reshape((lx, ly) + a.shape[1:])

        inner = (slice(1, nx + 1), slice(1, ly - 1))
        glob_inner = (slice(ox + 1, ox + nx + 1), slice(1, nyp - 1))
        for key in ("solid", "solid_id", "solid_boundary", "fluid_boundary",
                    "periodic_boundary", "surface_normals", "density",
                    "velocity", "pressure"):
            assert np.array_equal(local(getattr(f, key))[inner],
                                  grid(getattr(g, key))[glob_inner]), key
        # ghost columns carry the neighbour's solid flags where one exists
        left, right = neighbour_ranks(st.domain, st.boundary)
        rows = slice(1, ly - 1)
        if left is not None:
            src = (ox - 1) % (nxp - 2) + 1
            assert np.array_equal(local(f.solid)[0, rows], grid(g.solid)[src, rows])
        else:
            assert not local(f.solid)[0].any()
        if right is not None:
            src = (ox + nx) % (nxp - 2) + 1
            assert np.array_equal(local(f.solid)[nx + 1, rows],
                                  grid(g.solid)[src, rows])
        else:
            assert not local(f.solid)[nx + 1].any()
        # link lists: local padded index -> global padded index
        for n, el in enumerate(st.boundary.boundary_elements):
            ref = whole.boundary.boundary_elements[n]
            assert np.array_equal(el.out_list, ref.out_list)
            assert el.type_fluid == ref.type_fluid
            li, lj = np.divmod(el.boundary_nodes, ly)
            gl = (li + ox) * nyp + lj
            assert set(gl.tolist()) <= set(ref.boundary_nodes.tolist())
            assert len(set(gl.tolist())) == len(gl)
            owned_nodes[n] |= set(gl.tolist())
    for n, ref in enumerate(whole.boundary.boundary_elements):
        assert owned_nodes[n] == set(ref.boundary_nodes.tolist())
