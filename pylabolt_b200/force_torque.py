"""Momentum-exchange forces on walls and obstacles: the b200 mirror of
BoundaryOperator.compute_force (pylabolt/base/boundary_operator.py:204-237),
ObstacleOperator.compute_force_torque (base/obstacle_operator.py:75-109) and
their kernels (pylabolt/parallel/cpu/force_torque_kernels.py).

The reference sums, over the links that end on a wall / solid node,
``pop[i,k] c_k - pop_new[i,inv k] c_inv(k)`` with pop = post-collision and
pop_new = post-stream populations of the same step.  In the fused step both
values exist only in the registers of the link-list kernel, which (when asked,
PLB_RECORD_LINKS) writes their sum per link; the O(perimeter) reduction per
wall element / obstacle is done here on the host, in a fixed order.

Deviation, for cost: the reference reduces the obstacle force on every step
(single_time_step phase 9); static bodies never use it, so here it is
evaluated on the steps on which it is written or asked for.
"""
import numpy as np

_CX = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1], dtype=np.int64)
_CY = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1], dtype=np.int64)
_INV = np.array([0, 3, 4, 1, 2, 7, 8, 5, 6], dtype=np.int64)


class MomentumExchange:
    def __init__(self, state, link_inds, every_element=False):
        """``link_inds``: padded flat indices of libplb's link nodes, in list
        order (plb_link_nodes).  ``every_element``: also reduce the elements
        that are not walls (the reference's kernel can be called on any
        element; its operator only calls it on walls) -- parity tests."""
        self.state = state
        self.link_inds = np.asarray(link_inds, dtype=np.int64)
        self.n_links = self.link_inds.shape[0]
        order = np.argsort(self.link_inds, kind="stable")
        sorted_inds = self.link_inds[order]

        def rows_of(nodes):
            """list position of each node, -1 if it is not a link node."""
            nodes = np.asarray(nodes, dtype=np.int64)
            if self.n_links == 0 or nodes.size == 0:
                return np.full(nodes.shape, -1, dtype=np.int64)
            pos = np.clip(np.searchsorted(sorted_inds, nodes), 0,
                          self.n_links - 1)
            hit = sorted_inds[pos] == nodes
            return np.where(hit, order[pos], -1)

        f = state.fields
        # walls: three links per node of every boundary element
        self.wall_terms = []
        for el in state.boundary.boundary_elements:
            # only wall elements are reduced (boundary_operator.py:224-230);
            # the rows of inlets / outlets / periodic pairs stay zero
            nodes = (el.boundary_nodes[~f.solid[el.boundary_nodes]]
                     if (el.wall or every_element)
                     else np.zeros(0, dtype=np.int64))
            rows = rows_of(nodes)
            rows = rows[rows >= 0]
            out = np.asarray(el.out_list, dtype=np.int64)
            self.wall_terms.append(
                (np.repeat(rows, 3), np.tile(out - 1, rows.size),
                 np.tile(_CX[out], rows.size).astype(np.float64),
                 np.tile(_CY[out], rows.size).astype(np.float64)))
        # obstacles: links of fluid boundary nodes into solid nodes
        nyp = int(state.domain.shape[1])
        self.body_terms = []
        for body in state.obstacle.obstacles:
            nodes = np.flatnonzero(f.fluid_boundary & ~f.ghost_node &
                                   (f.solid_id == body.id))
            rows = rows_of(nodes)
            keep = rows >= 0
            nodes, rows = nodes[keep], rows[keep]
            x, y = np.divmod(nodes, nyp)
            i_glob = x - 1 + int(state.domain.offset[0])
            j_glob = y - 1 + int(state.domain.offset[1])
            ref = _RefPoint(body.ref_point)
            rx, ry = ref.min_image(i_glob, j_glob, state.mesh.grid_global_shape,
                                   state.boundary.x_periodic,
                                   state.boundary.y_periodic)
            r_list, q_list, cx_l, cy_l, rx_l, ry_l = [], [], [], [], [], []
            for k in range(1, 9):
                nb = (x + _CX[k]) * nyp + (y + _CY[k])
                sel = f.solid[nb]
                r_list.append(rows[sel])
                q_list.append(np.full(int(sel.sum()), k - 1, dtype=np.int64))
                cx_l.append(np.full(int(sel.sum()), float(_CX[k])))
                cy_l.append(np.full(int(sel.sum()), float(_CY[k])))
                rx_l.append(np.asarray(rx, dtype=np.float64)[sel])
                ry_l.append(np.asarray(ry, dtype=np.float64)[sel])
            self.body_terms.append(tuple(np.concatenate(v) for v in
                                         (r_list, q_list, cx_l, cy_l, rx_l,
                                          ry_l)))

    # ------------------------------------------------------------------
    def boundary_forces(self, exchange):
        """(n_elements, 2) local force per boundary element."""
        out = np.zeros((len(self.wall_terms), 2), dtype=np.float64)
        for n, (rows, cols, cx, cy) in enumerate(self.wall_terms):
            if rows.size:
                value = exchange[rows, cols]
                out[n, 0] = np.sum(value * cx)
                out[n, 1] = np.sum(value * cy)
        return out

    def obstacle_forces(self, exchange):
        """(n_obstacles, 3) local (force_x, force_y, torque) per obstacle."""
        out = np.zeros((len(self.body_terms), 3), dtype=np.float64)
        for n, (rows, cols, cx, cy, rx, ry) in enumerate(self.body_terms):
            if rows.size:
                value = exchange[rows, cols]
                vx, vy = value * cx, value * cy
                out[n, 0] = np.sum(vx)
                out[n, 1] = np.sum(vy)
                out[n, 2] = np.sum(rx * vy - ry * vx)
        return out

    def initial_exchange(self, lattice):
        """pop = pop_new = f_eq(rho0, u0) before the first step
        (Solver.run computes the forces once at time_step 0,
        solvers/fluidLB.py:324-337): f_eq,k + f_eq,inv(k) per link."""
        f = self.state.fields
        rho = f.density[self.link_inds]
        ux = f.velocity[self.link_inds, 0]
        uy = f.velocity[self.link_inds, 1]
        u2 = ux * ux + uy * uy
        feq = np.zeros((self.n_links, 9), dtype=np.float64)
        for k in range(9):
            cu = lattice.cx[k] * ux + lattice.cy[k] * uy
            feq[:, k] = lattice.weights[k] * rho * (
                1 + lattice.inv_cs_2 * cu + 0.5 * lattice.inv_cs_4 * cu * cu -
                0.5 * lattice.inv_cs_2 * u2)
        return feq[:, 1:] + feq[:, _INV[1:]]


class _RefPoint:
    """Minimum-image vector from a reference point
    (force_torque_kernels.py:52-66), same arithmetic as the obstacle code."""

    def __init__(self, point):
        self.center = np.asarray(point, dtype=np.float64)

    def min_image(self, i_glob, j_glob, grid, x_periodic, y_periodic):
        from .obstacle import Body
        return Body.min_image(self, i_glob, j_glob, grid, x_periodic,
                              y_periodic)
