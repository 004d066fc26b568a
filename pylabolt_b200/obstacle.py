"""Static obstacles: the b200 mirror of pylabolt/base/obstacle.py and of the
setup half of pylabolt/base/obstacle_operator.py.

The solid mask, solid ids, the rigid-body velocity / density stored on solid
nodes, the fluid / solid boundary flags and the surface normals are produced
with the same predicates and the same IEEE operations as the reference's numba
kernels (pylabolt/parallel/cpu/obstacle_kernels.py), but vectorised and
restricted to bounding boxes, so the flags are bit-identical and a 16384^2
lattice needs no per-node python loop.

Scope: bodies whose rasterised mask does not change in time -- static bodies
and circles with a prescribed rotation.  Moving bodies are refused: the
upstream path is unfinished (base/obstacle_operator.py:62-73).
"""
import math

import numpy as np

from .helpers import print_log

_CX = (0, 1, 0, -1, 0, 1, -1, -1, 1)
_CY = (0, 0, 1, 0, -1, 1, 1, -1, -1)


def _number(value, name, what, kind):
    if type(value) not in (float, int):
        raise ValueError(f"{what} must be float or int for obstacle type "
                         f"{kind}: {name}")
    return value


class Body:
    """Common part of Circle / Ellipse (base/obstacle.py:392-588, 591-825)."""

    type = None

    def __init__(self, obstacle_id, name, spec, control):
        self.id = obstacle_id
        self.name = name
        prec = control.precision
        self.force = np.zeros(2, dtype=prec)
        self.torque = prec(0)
        self.motion_type = None
        self.degree_of_freedom = None
        self.linear_velocity = np.zeros(2, dtype=prec)
        self.angular_velocity = prec(0)
        self.calculated = False
        self.rotation_allowed = False
        self.translation_allowed = False
        self.inclination_angle = prec(0)
        self.read_shape(spec, prec)
        for key, label in (("center", "center"), ("density", "density"),
                           ("static", "static")):
            if key not in spec:
                raise ValueError(f"{label} missing in obstacle: {name}")
        if not isinstance(spec["center"], list):
            raise ValueError(
                f"center must be list for obstacle type {self.type}: {name}")
        self.center = np.array(spec["center"], dtype=prec)
        self.ref_point = self.center
        self.solid_density = prec(_number(spec["density"], name, "density",
                                          self.type))
        if not isinstance(spec["static"], (bool, np.bool_)):
            raise ValueError(f"static must be True/False for obstacle type "
                             f"{self.type}: {name}")
        self.static = spec["static"]
        if not self.static:
            self.read_motion(spec, name, prec)

    def read_motion(self, spec, name, prec):
        if "solid_motion_dict" not in spec:
            raise ValueError("solid_motion_dict missing in obstacle: " + name)
        motion = spec["solid_motion_dict"]
        for key in ("type", "degree_of_freedom", "linear_velocity",
                    "angular_velocity"):
            if key not in motion:
                raise ValueError(key + " missing in solid_motion_dict: " + name)
        self.motion_type = motion["type"]
        if self.motion_type not in ("fixed_velocity", "calculated"):
            raise ValueError("Unsupported motion type: " + self.motion_type +
                             " in obstacle: " + name)
        self.degree_of_freedom = motion["degree_of_freedom"]
        if self.degree_of_freedom not in ("rotation", "translation", "both"):
            raise ValueError("Unsupported degree of freedom: " +
                             self.degree_of_freedom + " in obstacle: " + name)
        if not isinstance(motion["linear_velocity"], list):
            raise ValueError("linear_velocity must be list [ux, uy]: " + name)
        self.linear_velocity = np.array(motion["linear_velocity"], dtype=prec)
        if not isinstance(motion["angular_velocity"], (int, float)):
            raise ValueError("angular_velocity must be int/float: " + name)
        self.angular_velocity = np.array(motion["angular_velocity"], dtype=prec)
        self.calculated = self.motion_type == "calculated"
        self.rotation_allowed = self.degree_of_freedom in ("rotation", "both")
        self.translation_allowed = self.degree_of_freedom in ("translation",
                                                              "both")
        mask_is_invariant = (self.type == "circle" and not self.calculated
                             and not self.translation_allowed)
        if not mask_is_invariant:
            raise ValueError(
                "moving obstacles are not supported by the b200 back end "
                "(only static bodies and circles with a prescribed rotation): "
                + name)

    # -- geometry ------------------------------------------------------------
    def min_image(self, i_glob, j_glob, grid, x_periodic, y_periodic):
        """Centre-to-node vector with the minimum-image convention
        (is_circle / is_ellipse, cpu/obstacle_kernels.py:13-40, 149-186)."""
        nx_glob, ny_glob = int(grid[0]), int(grid[1])
        rx = i_glob - self.center[0]
        ry = j_glob - self.center[1]
        rx_min, ry_min = rx, ry
        if x_periodic:
            for shifted in (rx + nx_glob, rx - nx_glob):
                rx_min = np.where(np.abs(shifted) < np.abs(rx_min), shifted,
                                  rx_min)
        if y_periodic:
            for shifted in (ry + ny_glob, ry - ny_glob):
                ry_min = np.where(np.abs(shifted) < np.abs(ry_min), shifted,
                                  ry_min)
        return rx_min, ry_min


class Circle(Body):
    type = "circle"

    def read_shape(self, spec, prec):
        if "radius" not in spec:
            raise ValueError("radius missing in obstacle: " + self.name)
        self.radius = prec(_number(spec["radius"], self.name, "radius",
                                   "circle"))
        self.extent = float(self.radius)

    def finish(self):
        self.mass = np.pi * self.radius * self.radius * self.solid_density
        self.moment_of_inertia = self.mass * self.radius * self.radius / 2

    def inside(self, rx, ry):
        return rx * rx + ry * ry <= self.radius * self.radius

    def normal(self, rx, ry):
        mag = np.sqrt(rx * rx + ry * ry)
        return rx / mag, ry / mag

    @property
    def properties(self):
        return {"obstacle id": self.id, "obstacle name": self.name,
                "obstacle type": self.type, "density": self.solid_density,
                "center": [float(v) for v in self.center],
                "radius": self.radius, "static obstacle": self.static}


class Ellipse(Body):
    type = "ellipse"

    def read_shape(self, spec, prec):
        for key in ("semi_major_axis", "semi_minor_axis", "inclination_angle"):
            if key not in spec:
                raise ValueError(f"{key} missing in obstacle: {self.name}")
        self.semi_major_axis = prec(_number(
            spec["semi_major_axis"], self.name, "semi_major_axis", "ellipse"))
        self.semi_minor_axis = prec(_number(
            spec["semi_minor_axis"], self.name, "semi_minor_axis", "ellipse"))
        angle = _number(spec["inclination_angle"], self.name,
                        "inclination_angle", "ellipse representing angle in degrees")
        self.extent = float(max(self.semi_major_axis, self.semi_minor_axis))
        self._angle_deg = angle

    def finish(self):
        self.inclination_angle = type(self.semi_major_axis)(self._angle_deg) * \
            np.pi / 180
        # libm cos / sin, like the numba kernels (construct_ellipse :207-208)
        self.cos_alpha = math.cos(self.inclination_angle)
        self.sin_alpha = math.sin(self.inclination_angle)
        a, b = self.semi_major_axis, self.semi_minor_axis
        self.mass = np.pi * a * b * self.solid_density
        self.moment_of_inertia = self.mass * (a * a + b * b) / 4

    def _project(self, rx, ry):
        x_proj = rx * self.cos_alpha + ry * self.sin_alpha
        y_proj = -rx * self.sin_alpha + ry * self.cos_alpha
        return x_proj, y_proj

    def inside(self, rx, ry):
        x_proj, y_proj = self._project(rx, ry)
        a, b = self.semi_major_axis, self.semi_minor_axis
        return (x_proj * x_proj) / (a * a) + (y_proj * y_proj) / (b * b) <= 1

    def normal(self, rx, ry):
        """compute_normals_ellipse, cpu/obstacle_kernels.py:242-304."""
        x_proj, y_proj = self._project(rx, ry)
        gx = x_proj / (self.semi_major_axis * self.semi_major_axis)
        gy = y_proj / (self.semi_minor_axis * self.semi_minor_axis)
        x_g = gx * self.cos_alpha - gy * self.sin_alpha
        y_g = gx * self.sin_alpha + gy * self.cos_alpha
        mag = np.sqrt(x_g * x_g + y_g * y_g)
        return x_g / mag, y_g / mag

    @property
    def properties(self):
        return {"obstacle id": self.id, "obstacle name": self.name,
                "obstacle type": self.type, "density": self.solid_density,
                "center": [float(v) for v in self.center],
                "semi-major axis": self.semi_major_axis,
                "semi-minor axis": self.semi_minor_axis,
                "inclination angle": np.rad2deg(self.inclination_angle),
                "static obstacle": self.static}


def _runs(lo, hi, n, periodic, origin, count):
    """Local index runs [a, b) of a rank that owns global indices
    origin-1 .. origin+count (ghosts included) and that can see the global
    interval [lo, hi] of a body, periodic images included."""
    shifts = (-n, 0, n) if periodic else (0,)
    out = []
    for shift in shifts:
        a = max(lo + shift, origin - 1)
        b = min(hi + shift, origin + count)
        if a <= b:
            out.append((a - origin + 1, b - origin + 2))   # padded local
    return out


class Obstacle:
    """obstacle_dict -> solid mask and obstacle boundary flags
    (base/obstacle.py:10-300 + base/obstacle_operator.py:20-45)."""

    def __init__(self, simulation, mesh, domain, control, fields, boundary,
                 fluid=False, phase=False, scalar=False, verbose=True):
        # signature of base/obstacle.py:11-23
        self.fluid, self.phase, self.scalar = fluid, phase, scalar
        rank = domain.mpi_rank
        print_log("-" * 80, rank, verbose)
        print_log("Setting up obstacles...\n", rank, verbose)
        if not hasattr(simulation, "obstacle_dict"):
            raise ValueError("obstacle_dict not found in simulation.py file")
        self.obstacle_dict = simulation.obstacle_dict
        self.compute_force_torque = False
        self.ref_point_torque = np.zeros(2, dtype=control.precision)
        self.write_obstacle_data = False
        self.write_interval = 1
        self.obstacles = []
        if "options" not in self.obstacle_dict:
            raise ValueError("options missing in obstacle_dict")
        self.read_options_dict(self.obstacle_dict["options"], domain, control,
                               verbose)
        names = [key for key in self.obstacle_dict if key != "options"]
        for obstacle_id, name in enumerate(names):
            spec = self.obstacle_dict[name]
            if "type" not in spec:
                raise ValueError("type missing in obstacle: " + name)
            if spec["type"] == "circle":
                body = Circle(obstacle_id, name, spec, control)
            elif spec["type"] == "ellipse":
                body = Ellipse(obstacle_id, name, spec, control)
            else:
                raise ValueError("Unsupported obstacle type: " + spec["type"])
            body.finish()
            self.rasterise(body, mesh, domain, fields, boundary)
            self.obstacles.append(body)
            for key, value in body.properties.items():
                print_log(f"{key:20s}: {value}", rank, verbose)
        self.no_of_obstacles = len(self.obstacles)
        self.all_obstacles_static = all(b.static for b in self.obstacles)
        if not self.all_obstacles_static and not self.compute_force_torque:
            raise ValueError("If any obstacle is not static, "
                             "then compute_force_torque must be True")
        self.mark_boundary_nodes(mesh, domain, fields, boundary)
        print_log("Setting up obstacles done!", rank, verbose)
        print_log("-" * 80, rank, verbose)

    def read_options_dict(self, options, domain, control, verbose):
        rank = domain.mpi_rank
        if "compute_force_torque" in options:
            self.compute_force_torque = options["compute_force_torque"]
            if not isinstance(self.compute_force_torque, (bool, np.bool_)):
                raise ValueError("compute_force_torque must be bool:" +
                                 " True/False (default: False)")
        if "ref_point_torque" in options:
            if self.compute_force_torque is False:
                print_log("WARNING! ref_point_torque ignored in "
                          "obstacle_dict options as compute_force_torque is "
                          "False", rank, verbose)
            else:
                if not isinstance(options["ref_point_torque"], list):
                    raise ValueError("ref_point_torque must be list: (x, y)")
                self.ref_point_torque = np.array(options["ref_point_torque"],
                                                 dtype=control.precision)
        if "write_obstacle_data" in options:
            self.write_obstacle_data = True
            sub = options["write_obstacle_data"]
            if "interval" not in sub:
                raise ValueError("interval missing obstacle_dict options!")
            self.write_interval = sub["interval"]
            if not isinstance(self.write_interval, (int, type(None))):
                raise ValueError(
                    "interval must be int or None in obstacle_dict options")
            if isinstance(self.write_interval, int) and self.write_interval <= 0:
                raise ValueError("if interval is int then it must be > 0"
                                 " in obstacle_dict options")
        print_log("compute_force_torque: " + str(self.compute_force_torque),
                  rank, verbose)

    # -- solid mask ----------------------------------------------------------
    def rasterise(self, body, mesh, domain, fields, boundary):
        """construct_circle / construct_ellipse
        (cpu/obstacle_kernels.py:43-93, 189-239) on this rank's nodes, plus
        the ghost ring where a neighbour rank or a periodic image exists --
        what the reference obtains by halo-exchanging ``solid`` and
        ``solid_id`` (base/obstacle_operator.py:36-41).  Only interior solid
        nodes receive the body's density and rigid-body velocity."""
        grid = mesh.grid_global_shape
        nxp, nyp = int(domain.shape[0]), int(domain.shape[1])
        reach = int(math.ceil(body.extent)) + 2
        lo_x = int(math.floor(body.center[0])) - reach
        hi_x = int(math.ceil(body.center[0])) + reach
        lo_y = int(math.floor(body.center[1])) - reach
        hi_y = int(math.ceil(body.center[1])) + reach
        solid = fields.solid.reshape(nxp, nyp)
        solid_id = fields.solid_id.reshape(nxp, nyp)
        density = fields.density.reshape(nxp, nyp)
        velocity = fields.velocity.reshape(nxp, nyp, 2)
        for a_x, b_x in _runs(lo_x, hi_x, int(grid[0]), boundary.x_periodic,
                              int(domain.offset[0]), nxp - 2):
            for a_y, b_y in _runs(lo_y, hi_y, int(grid[1]), boundary.y_periodic,
                                  int(domain.offset[1]), nyp - 2):
                i_loc = np.arange(a_x, b_x, dtype=np.int64)[:, None]
                j_loc = np.arange(a_y, b_y, dtype=np.int64)[None, :]
                i_glob = i_loc - 1 + int(domain.offset[0])
                j_glob = j_loc - 1 + int(domain.offset[1])
                ghost = ((i_loc == 0) | (i_loc == nxp - 1) |
                         (j_loc == 0) | (j_loc == nyp - 1))
                # a ghost node mirrors the global node it stands for, if any
                i_img, j_img = i_glob, j_glob
                exists = np.ones(ghost.shape, dtype=bool) & True
                if boundary.x_periodic:
                    i_img = i_glob % int(grid[0])
                else:
                    exists = exists & (i_glob >= 0) & (i_glob < int(grid[0]))
                if boundary.y_periodic:
                    j_img = j_glob % int(grid[1])
                else:
                    exists = exists & (j_glob >= 0) & (j_glob < int(grid[1]))
                i_eval = np.where(ghost, i_img, i_glob) + 0 * j_glob
                j_eval = np.where(ghost, j_img, j_glob) + 0 * i_glob
                rx, ry = body.min_image(i_eval, j_eval, grid,
                                        boundary.x_periodic, boundary.y_periodic)
                hit = body.inside(rx, ry) & exists
                block = (slice(a_x, b_x), slice(a_y, b_y))
                solid[block] |= hit
                solid_id[block] = np.where(hit, body.id, solid_id[block])
                own = hit & ~ghost
                density[block] = np.where(own, body.solid_density, density[block])
                # The reference never exchanges `velocity`, so a solid node in
                # the ghost ring has zero wall velocity: true of a periodic
                # image on one rank (kept, it is what the golden runs contain)
                # but, on several ranks, it would make the moving-wall term
                # depend on where the slabs are cut.  Ghost nodes that stand
                # for a node of the neighbouring slab therefore carry that
                # node's rigid-body velocity: results do not depend on the
                # decomposition and equal the single-rank reference.
                in_domain = ((i_glob >= 0) & (i_glob < int(grid[0])) &
                             (j_glob >= 0) & (j_glob < int(grid[1])))
                moving = hit & (~ghost | in_domain)
                vel = velocity[block]
                vel[..., 0] = np.where(
                    moving, body.linear_velocity[0] - body.angular_velocity * ry,
                    vel[..., 0])
                vel[..., 1] = np.where(
                    moving, body.linear_velocity[1] + body.angular_velocity * rx,
                    vel[..., 1])

    # -- obstacle boundary nodes and normals -----------------------------------
    def mark_boundary_nodes(self, mesh, domain, fields, boundary):
        """compute_obstacle_boundary, check_fluid_boundary_overlap and
        compute_normals_* (cpu/obstacle_kernels.py:307-384, 96-143, 242-304),
        restricted to the bounding box of the solid nodes."""
        nxp, nyp = int(domain.shape[0]), int(domain.shape[1])
        solid = fields.solid.reshape(nxp, nyp)
        rows = np.flatnonzero(solid.any(axis=1))
        if rows.size == 0:
            return
        cols = np.flatnonzero(solid.any(axis=0))
        x0, x1 = max(1, rows[0] - 1), min(nxp - 2, rows[-1] + 1)
        y0, y1 = max(1, cols[0] - 1), min(nyp - 2, cols[-1] + 1)
        if x0 > x1 or y0 > y1:
            return
        box = (slice(x0, x1 + 1), slice(y0, y1 + 1))
        solid_id = fields.solid_id.reshape(nxp, nyp)
        is_solid = solid[box]
        touches_fluid = np.zeros_like(is_solid)
        touches_solid = np.zeros_like(is_solid)
        first_id = np.full(is_solid.shape, -1, dtype=solid_id.dtype)
        for k in range(8, 0, -1):       # lowest k wins, like the `break`
            nb = (slice(x0 + _CX[k], x1 + 1 + _CX[k]),
                  slice(y0 + _CY[k], y1 + 1 + _CY[k]))
            touches_fluid |= ~solid[nb]
            touches_solid |= solid[nb]
            first_id = np.where(solid[nb], solid_id[nb], first_id)
        solid_boundary = fields.solid_boundary.reshape(nxp, nyp)
        fluid_boundary = fields.fluid_boundary.reshape(nxp, nyp)
        solid_boundary[box] |= is_solid & touches_fluid
        new_fluid_boundary = ~is_solid & touches_solid
        fluid_boundary[box] |= new_fluid_boundary
        solid_id[box] = np.where(new_fluid_boundary, first_id, solid_id[box])

        own_id = solid_id[box]
        overlap = np.zeros_like(is_solid)
        for k in range(1, 9):
            nb = (slice(x0 + _CX[k], x1 + 1 + _CX[k]),
                  slice(y0 + _CY[k], y1 + 1 + _CY[k]))
            overlap |= ~((solid_id[nb] == own_id) | (solid_id[nb] == -1))
        n_overlap = int(np.count_nonzero(overlap & fluid_boundary[box]))
        self.local_fluid_boundary_overlap = n_overlap

        if not hasattr(fields, "surface_normals"):
            return              # a geometry-only container (reference tests)
        normals = fields.surface_normals.reshape(nxp, nyp, 2)
        surface = (solid_boundary[box] | fluid_boundary[box])
        i_glob = (np.arange(x0, x1 + 1, dtype=np.int64)[:, None] - 1 +
                  int(domain.offset[0])) + 0 * np.arange(y0, y1 + 1)[None, :]
        j_glob = (np.arange(y0, y1 + 1, dtype=np.int64)[None, :] - 1 +
                  int(domain.offset[1])) + 0 * np.arange(x0, x1 + 1)[:, None]
        for body in self.obstacles:
            sel = surface & (own_id == body.id)
            if not sel.any():
                continue
            rx, ry = body.min_image(i_glob[sel], j_glob[sel],
                                    mesh.grid_global_shape,
                                    boundary.x_periodic, boundary.y_periodic)
            n_x, n_y = body.normal(rx, ry)
            sub = normals[box]
            sub[sel, 0] = n_x
            sub[sel, 1] = n_y

    def check_overlap(self, comm):
        """Two bodies may not share a fluid boundary node
        (base/obstacle_operator.py:160-182)."""
        import numpy as _np
        local = _np.array([getattr(self, "local_fluid_boundary_overlap", 0)],
                          dtype=_np.int64)
        total = _np.zeros_like(local)
        comm.Allreduce(local, total)
        if total[0] > 0:
            raise RuntimeError(
                f"Fluid Boundary node overlap detected for {int(total[0])} "
                "nodes!\nThis indicates two solid obstacles have a common "
                "fluid boundary node which is illegal!\nTo avoid this issue, "
                "ensure solid particle surfaces have 2-3 lattice nodes in "
                "between.")
