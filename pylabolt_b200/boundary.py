"""Domain-edge boundaries: the b200 mirror of pylabolt/base/boundary.py.

``Boundary`` parses ``boundary_dict`` with the reference's keywords and error
messages and produces one ``BoundaryElement`` per segment, in dict order,
whose link lists (boundary_nodes, out_list, inv_list, surface_normals) are
bit-identical to the reference's (pinned by tests/golden).  Two additions:
the ``zero_gradient`` type (listed by the reference's README.rst:83 but never
implemented upstream) and the README-era camelCase spellings as aliases.
"""
import numpy as np

from .helpers import print_log

# location -> (inward normal, outgoing directions, incoming directions)
# pylabolt/base/boundary.py:63-78
EDGE_TABLE = {
    "bottom": ((0, 1), (4, 7, 8), (2, 5, 6)),
    "top": ((0, -1), (2, 5, 6), (4, 7, 8)),
    "left": ((1, 0), (3, 6, 7), (1, 8, 5)),
    "right": ((-1, 0), (1, 8, 5), (3, 6, 7)),
}

SUPPORTED_FLUID_BCS = ["bounce_back", "fixed_velocity", "fixed_pressure",
                       "periodic", "zero_gradient"]
# legacy 0.1.x spellings (docs/boundary_conditions.rst) accepted as aliases
LEGACY_ALIASES = {"bounceBack": "bounce_back", "fixedU": "fixed_velocity",
                  "fixedPressure": "fixed_pressure",
                  "zeroGradient": "zero_gradient"}


class BoundaryElement:
    """One axis-aligned segment on a domain edge (base/boundary.py:7-127)."""

    def __init__(self, boundary_name, segment, segment_no, orientation,
                 location, domain, control, fields, wall=False,
                 fluid_config=None):
        self.name = boundary_name + "_" + str(int(segment_no))
        self.orientation = orientation
        self.location = location
        self.segment = np.array(segment)
        self.wall = wall
        self.force = np.zeros(2, dtype=control.precision)
        self.type_fluid = fluid_config["type"]
        self.periodic = self.type_fluid == "periodic"
        self.scalar_fluid = control.precision(
            0 if fluid_config["scalar_value"] is None
            else fluid_config["scalar_value"])
        self.vector_fluid = (
            np.zeros(2, dtype=control.precision)
            if fluid_config["vector_value"] is None
            else np.array(fluid_config["vector_value"], dtype=control.precision))
        normal, out_list, inv_list = EDGE_TABLE[location]
        # the reference stores the normal as floats (boundary.py:64-76); the
        # integer copy is what indexes nodes (SURVEY.md 8(c) shim 4)
        self.surface_normals = np.array(normal, dtype=control.precision)
        self.normal = np.array(normal, dtype=np.int64)
        self.out_list = np.array(out_list, dtype=int)
        self.inv_list = np.array(inv_list, dtype=int)
        self.boundary_nodes = self.allocate_boundary_nodes(segment, domain)
        if self.periodic:
            fields.periodic_boundary[self.boundary_nodes] = True

    def allocate_boundary_nodes(self, segment, domain):
        """Padded flat indices of the segment's nodes owned by this rank, in
        increasing global coordinate (base/boundary.py:86-127)."""
        (x_min, y_min), (x_max, y_max) = segment
        if self.orientation == "horizontal":
            i_glob = np.arange(x_min, x_max + 1, dtype=np.int64)
            j_glob = np.full_like(i_glob, y_min)
        else:
            j_glob = np.arange(y_min, y_max + 1, dtype=np.int64)
            i_glob = np.full_like(j_glob, x_min)
        i = i_glob - int(domain.offset[0])
        j = j_glob - int(domain.offset[1])
        owned = ((i >= 0) & (j >= 0) & (i < domain.shape[0] - 2) &
                 (j < domain.shape[1] - 2))
        return ((i[owned] + 1) * int(domain.shape[1]) + (j[owned] + 1)).astype(int)


class Boundary:
    """boundary_dict -> boundary elements (base/boundary.py:130-653)."""

    def __init__(self, simulation, mesh, domain, control, fields, fluid=False,
                 phase=False, scalar=False, verbose=True):
        # signature of base/boundary.py:131-142
        rank = domain.mpi_rank
        print_log("-" * 80, rank, verbose)
        print_log("Setting up domain boundaries...\n", rank, verbose)
        if not hasattr(simulation, "boundary_dict"):
            raise ValueError("boundary_dict not found in simulation.py file")
        self.fluid, self.phase, self.scalar = fluid, phase, scalar
        self.boundary_dict = simulation.boundary_dict
        self.compute_force = False
        self.write_boundary_data = False
        self.write_interval = 1
        self.x_periodic = False
        self.y_periodic = False
        self.boundary_elements = []
        if "options" not in self.boundary_dict:
            raise ValueError("options missing in boundary_dict")
        self.read_options_dict(self.boundary_dict["options"], domain, verbose)
        for name, user_dict in self.boundary_dict.items():
            if name == "options":
                continue
            self.read_user_boundary_dict(name, user_dict, mesh, domain,
                                         control, fields, verbose)
        self.no_of_boundaries = len(self.boundary_elements)
        self.local_force = np.zeros((self.no_of_boundaries, 2),
                                    dtype=control.precision)
        self.global_force = np.zeros_like(self.local_force)
        print_log("Setting up domain boundaries done!", rank, verbose)
        print_log("-" * 80, rank, verbose)

    def read_options_dict(self, options, domain, verbose):
        rank = domain.mpi_rank
        print_log("Setting boundary options", rank, verbose)
        if "compute_force" in options:
            self.compute_force = options["compute_force"]
            if not isinstance(self.compute_force, (bool, np.bool_)):
                raise ValueError("compute_force must be a bool:" +
                                 " True/False (default: False)")
        print_log("compute_force: " + str(self.compute_force), rank, verbose)
        if "write_boundary_data" in options:
            self.write_boundary_data = True
            sub = options["write_boundary_data"]
            if "interval" not in sub:
                raise ValueError("interval missing boundary_dict options!")
            self.write_interval = sub["interval"]
            if not isinstance(self.write_interval, (int, type(None))):
                raise ValueError(
                    "interval must be int or None in boundary_dict options")
            if isinstance(self.write_interval, int) and self.write_interval <= 0:
                raise ValueError("if interval is int then it must be > 0"
                                 " in boundary_dict options")
        print_log("write_boundary_data: " + str(self.write_boundary_data),
                  rank, verbose)
        print_log("Boundary options set\n", rank, verbose)

    @staticmethod
    def _check_segments(name, segments):
        if not isinstance(segments, list):
            raise ValueError("segments must be a list object: " + name)
        if len(segments) == 0:
            raise ValueError("segments cannot be an empty list: " + name)
        for segment in segments:
            well_formed = (isinstance(segment, list) and len(segment) == 2 and
                           all(isinstance(pt, list) and len(pt) == 2
                               for pt in segment))
            if not well_formed:
                raise ValueError(
                    "segment must have structure [[x1, y1], [x2, y2]]: " + name)
            (x1, y1), (x2, y2) = segment
            if not (x2 >= x1 and y2 >= y1):
                raise ValueError(
                    "segment must satisfy x2 >= x1 and y2 >= y1: " + name)
            if not (x1 == x2 or y1 == y2):
                raise ValueError("segment must be axis-aligned "
                                 "(horizontal or vertical): " + name)

    @staticmethod
    def _fluid_config(name, user_dict, control):
        if "fluid" not in user_dict:
            raise ValueError("fluid missing in boundary: " + name)
        fluid = user_dict["fluid"]
        if "type" not in fluid:
            raise ValueError("type missing in fluid section of boundary: " + name)
        bc = LEGACY_ALIASES.get(fluid["type"], fluid["type"])
        if bc not in SUPPORTED_FLUID_BCS:
            raise ValueError(
                "Unsupported boundary condition for fluid: " + str(bc))
        config = {"type": bc, "scalar_value": None, "vector_value": None}
        if bc in ("fixed_velocity", "fixed_pressure"):
            if "value" not in fluid:
                raise ValueError(
                    "value missing in fluid section of boundary: " + name)
            value = fluid["value"]
            if bc == "fixed_velocity":
                if not isinstance(value, list) or len(value) != 2:
                    raise ValueError("value for fixed_velocity must be a "
                                     "list (ux, uy): " + name)
                config["vector_value"] = np.array(value, dtype=control.precision)
            else:
                if not isinstance(value, (int, float)):
                    raise ValueError("value for fixed_pressure must be a "
                                     "float or int: " + name)
                config["scalar_value"] = control.precision(value)
        return config

    def read_user_boundary_dict(self, name, user_dict, mesh, domain, control,
                                fields, verbose=True):
        if "wall" not in user_dict:
            raise ValueError("wall missing in boundary: " + name)
        wall = user_dict["wall"]
        if not isinstance(wall, (bool, np.bool_)):
            raise ValueError("wall entry must be True/False: " + name)
        if "segments" not in user_dict:
            raise ValueError("segments missing in boundary: " + name)
        segments = user_dict["segments"]
        self._check_segments(name, segments)
        config = self._fluid_config(name, user_dict, control)
        if config["type"] == "periodic":
            self.validate_periodic_boundary(name, segments, mesh)
        nx_glob, ny_glob = (int(v) for v in mesh.grid_global_shape)
        for segment_no, segment in enumerate(segments):
            (x1, y1), (x2, y2) = segment
            location = None
            if x2 - x1 == 0:
                orientation = "vertical"
                if x2 == nx_glob - 1:
                    location = "right"
                elif x2 == 0:
                    location = "left"
            else:
                orientation = "horizontal"
                if y2 == ny_glob - 1:
                    location = "top"
                elif y2 == 0:
                    location = "bottom"
            if location is None:
                raise ValueError(
                    "segment does not lie on a domain edge: " + name)
            self.boundary_elements.append(BoundaryElement(
                name, segment, segment_no, orientation, location, domain,
                control, fields, wall=wall, fluid_config=config))
            print_log(f"boundary {name} segment_{segment_no + 1}: {segment} "
                      f"| {orientation} | {location} | {config['type']}",
                      domain.mpi_rank, verbose)

    def validate_periodic_boundary(self, name, segments, mesh):
        """A periodic boundary is a pair of opposite, full-length edges
        (base/boundary.py:495-571)."""
        if len(segments) != 2:
            raise ValueError("For periodic boundary, each segment should be" +
                             " followed by a corresponding periodic pair: " + name)
        first, second = segments
        nx_glob, ny_glob = (int(v) for v in mesh.grid_global_shape)
        horizontal = first[1][1] - first[0][1] == 0
        vertical = first[1][0] - first[0][0] == 0
        if horizontal and second[1][1] - second[0][1] != 0:
            raise ValueError(
                "invalid periodic pair - different orientation: " + name)
        if vertical and second[1][0] - second[0][0] != 0:
            raise ValueError(
                "invalid periodic pair - different orientation: " + name)
        if horizontal:
            spans = [(seg[0][0], seg[1][0]) for seg in segments]
            if any(span != (0, nx_glob - 1) for span in spans):
                raise ValueError("horizontal periodic boundaries must span" +
                                 " the entire x-direction: " + name)
            if sorted([first[0][1], second[0][1]]) != [0, ny_glob - 1]:
                raise ValueError("horizontal periodic boundaries must connect" +
                                 " top-bottom boundaries: " + name)
            self.y_periodic = True
        elif vertical:
            spans = [(seg[0][1], seg[1][1]) for seg in segments]
            if any(span != (0, ny_glob - 1) for span in spans):
                raise ValueError("vertical periodic boundaries must span" +
                                 " the entire y-direction: " + name)
            if sorted([first[0][0], second[0][0]]) != [0, nx_glob - 1]:
                raise ValueError("vertical periodic boundaries must connect" +
                                 " left-right boundaries: " + name)
            self.x_periodic = True
