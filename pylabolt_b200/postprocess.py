"""Post-processing of the solver's output: ``--reconstruct`` and ``--to_vtk``.

Native equivalents of pylabolt/utils/reconstruct.py and
pylabolt/utils/npz2vtk.py, working on the same files (``metadata.json``,
``procs/proc_<r>/rank_metadata.json`` + ``t_<n>.npz``,
``output/fields/t_<n>.npz``) and producing the same results:

* reconstruct: the per-rank blocks are placed at their ``offset`` in global
  x-major arrays and saved to ``output/fields/t_<n>.npz``
  (reconstruct.py:184-260);
* to_vtk: ``output/vtk/t_<n>.vtk``, a legacy-format (4.2) RECTILINEAR_GRID
  with unit spacing and, as point data, ``point_ID`` plus every saved field
  (npz2vtk.py:119-208).  The reference goes through the ``vtk`` python
  package; the file is written directly here (no dependency), in the point
  order that writer uses (x fastest).

Neither needs a GPU.
"""
import json
import os
import re

import numpy as np

from .helpers import print_log


def _load_metadata(root_dir):
    path = os.path.join(root_dir, "metadata.json")
    if not os.path.isfile(path):
        raise FileNotFoundError("metadata.json not found in " +
                                os.path.abspath(root_dir))
    with open(path) as f:
        meta = json.load(f)
    try:
        meta["mesh"]["size"], meta["mesh"]["shape"]
        meta["control"]["start_time"], meta["control"]["end_time"]
        meta["control"]["save_interval"]
        meta["decomposition"]["nx"], meta["decomposition"]["ny"]
        meta["fields_saved"]
    except KeyError as e:
        raise KeyError("Invalid metadata.json! missing key: " + str(e))
    return meta


def _empty_fields(meta):
    size = int(meta["mesh"]["size"])
    fields = {}
    for name, spec in meta["fields_saved"].items():
        shape = (size,) if spec["components"] == 1 else \
            (size, int(spec["components"]))
        fields[name] = np.zeros(shape, dtype=spec["dtype"])
    return fields


def _saved_times(directory):
    pattern = re.compile(r"^t_(\d+)\.npz$")
    times = [int(m.group(1)) for m in map(pattern.match, os.listdir(directory))
             if m]
    return sorted(times)


# ---------------------------------------------------------------------------
# --reconstruct
# ---------------------------------------------------------------------------
class ReconstructOperator:
    """reconstruct.py:66-300."""

    def __init__(self, metadata, root_dir=".", verbose=True):
        self.metadata = metadata
        self.root_dir = root_dir
        self.mesh_shape = tuple(int(v) for v in metadata["mesh"]["shape"])
        self.n_ranks = int(metadata["decomposition"]["nx"]) * \
            int(metadata["decomposition"]["ny"])
        self.fields = _empty_fields(metadata)
        print_log("\nFields to reconstruct:", 0, verbose)
        for name, field in self.fields.items():
            print_log(f"{name:<10}: {str(field.dtype):<5}", 0, verbose)
        self.domains = []
        for rank in range(self.n_ranks):
            directory = self._proc_dir(rank)
            if not os.path.isdir(directory):
                raise FileNotFoundError(
                    "processor directory not found in procs/ for rank: " +
                    str(rank))
            path = os.path.join(directory, "rank_metadata.json")
            if not os.path.isfile(path):
                raise FileNotFoundError(
                    "rank_metadata.json file not found for rank: " + str(rank))
            with open(path) as f:
                rank_meta = json.load(f)
            try:
                self.domains.append({
                    "shape": tuple(int(v) for v in rank_meta["domain_shape"]),
                    "offset": tuple(int(v) for v in rank_meta["offset"]),
                    "size": int(rank_meta["domain_size"])})
            except KeyError as e:
                raise KeyError("Invalid rank_metadata.json! missing key: " +
                               str(e) + " || rank: " + str(rank))

    def _proc_dir(self, rank):
        return os.path.join(self.root_dir, "procs", "proc_" + str(rank))

    def reconstruct_time(self, time_step, verbose=True):
        nx_g, ny_g = self.mesh_shape
        for rank, domain in enumerate(self.domains):
            path = os.path.join(self._proc_dir(rank),
                                "t_" + str(time_step) + ".npz")
            if not os.path.isfile(path):
                raise FileNotFoundError(
                    "output file not found for time step: " + str(time_step) +
                    " || rank: " + str(rank))
            local = np.load(path)
            (nx, ny), (ox, oy) = domain["shape"], domain["offset"]
            for name, field in self.fields.items():
                if name not in local.files:
                    print_log(f"{'missing field':<10}: {name:<20}"
                              f"{'time':<10}: {time_step:<20}"
                              f"{'rank':<10}: {rank:<20}", 0, True)
                    continue
                tail = field.shape[1:]
                block = local[name].reshape((nx, ny) + tail)
                field.reshape((nx_g, ny_g) + tail)[ox:ox + nx,
                                                   oy:oy + ny] = block
        out_dir = os.path.join(self.root_dir, "output", "fields")
        os.makedirs(out_dir, exist_ok=True)
        np.savez(os.path.join(out_dir, "t_" + str(time_step) + ".npz"),
                 **self.fields)
        print_log(f"{'Reconstruction done, time':<25}: {time_step:<5}", 0,
                  verbose)

    def reconstruct_all(self, verbose=True):
        c = self.metadata["control"]
        if c["save_interval"] is None:
            return []
        # the saved steps are start_time + k * save_interval by construction
        # (reconstruct.py:283-289); keep those that rank 0 actually wrote
        written = set(_saved_times(self._proc_dir(0)))
        times = [t for t in range(int(c["start_time"]), int(c["end_time"]) + 1)
                 if t % int(c["save_interval"]) == 0 and t in written]
        for t in times:
            self.reconstruct_time(t, verbose=verbose)
        return times


def reconstruct_data(option, time_step=0, verbose=True, root_dir="."):
    """reconstruct.py:303-335."""
    print_log("-" * 80, 0, verbose)
    print_log("Reconstructing fields...\n", 0, verbose)
    if option not in ("all", "time"):
        raise ValueError("Invalid reconstruction option!\n"
                         "Supported options: ['all', 'time']")
    op = ReconstructOperator(_load_metadata(root_dir), root_dir=root_dir,
                             verbose=verbose)
    if option == "all":
        op.reconstruct_all(verbose=verbose)
    else:
        op.reconstruct_time(int(time_step), verbose=verbose)
    print_log("\nReconstructing fields done!", 0, verbose)
    print_log("-" * 80, 0, verbose)


# ---------------------------------------------------------------------------
# --to_vtk
# ---------------------------------------------------------------------------
def _vtk_type(dtype):
    """npz2vtk.py:139-143: floats -> vtkDoubleArray, int64 / bool -> vtkIntArray."""
    kind = np.dtype(dtype).kind
    if kind == "f":
        return "double", ">f8", "%.17g"
    if kind in "iub":
        return "int", ">i4", "%d"
    raise ValueError("field dtype " + str(dtype) + " has no VTK array type")


def _write_array(out, name, values, components, binary):
    type_name, be, fmt = _vtk_type(values.dtype)
    flat = values.reshape(-1, components)
    out.write(f"{name} {components} {flat.shape[0]} {type_name}\n".encode())
    if binary:
        out.write(flat.astype(be).tobytes())
        out.write(b"\n")
    else:
        cast = flat.astype(np.float64 if type_name == "double" else np.int64)
        np.savetxt(out, cast, fmt=fmt)       # one tuple per line


class VTKOperator:
    """npz2vtk.py:16-250."""

    def __init__(self, metadata, root_dir=".", binary=False, verbose=True):
        self.metadata = metadata
        self.root_dir = root_dir
        self.binary = binary
        self.mesh_shape = tuple(int(v) for v in metadata["mesh"]["shape"])
        self.fields = _empty_fields(metadata)
        self.raw_data_path = os.path.join(root_dir, "output", "fields")
        self.vtk_save_path = os.path.join(root_dir, "output", "vtk")
        print_log("\nFields being converted to VTK:", 0, verbose)
        for name, field in self.fields.items():
            print_log(f"{name:<10}: {str(field.dtype):>5}", 0, verbose)

    def convert_time(self, time_step, verbose=True):
        path = os.path.join(self.raw_data_path, "t_" + str(time_step) + ".npz")
        if not os.path.isfile(path):
            raise FileNotFoundError("output file not found for time step: " +
                                    str(time_step))
        raw = np.load(path)
        data = {}
        for name, template in self.fields.items():
            if name not in raw.files:
                raise KeyError(f"{'missing field':<10}: {name:<20}"
                               f"{'time':<10}: {time_step:<20}")
            if raw[name].shape != template.shape:
                raise ValueError(
                    "Size and shape of raw data does not match with "
                    "metadata.json\nfield: " + name + " | raw data shape: " +
                    str(raw[name].shape) + " | metadata shape: " +
                    str(template.shape))
            if raw[name].dtype != template.dtype:
                raise ValueError(
                    "dtype of raw data does not match with metadata.json\n"
                    "field: " + name + " | raw data dtype: " +
                    str(raw[name].dtype) + " | metadata dtype: " +
                    str(template.dtype))
            data[name] = raw[name]
        self.write_vtk_fields(time_step, data, verbose=verbose)

    def write_vtk_fields(self, time_step, data, verbose=True):
        nx, ny = self.mesh_shape
        n = nx * ny
        os.makedirs(self.vtk_save_path, exist_ok=True)
        # VTK point order: x fastest; node (i, j) is stored at i * ny + j
        def point_order(a):
            tail = a.shape[1:]
            return np.swapaxes(a.reshape((nx, ny) + tail), 0, 1)
        path = os.path.join(self.vtk_save_path, "t_" + str(time_step) + ".vtk")
        with open(path, "wb") as out:
            out.write(b"# vtk DataFile Version 4.2\nvtk output\n")
            out.write(b"BINARY\n" if self.binary else b"ASCII\n")
            out.write(b"DATASET RECTILINEAR_GRID\n")
            out.write(f"DIMENSIONS {nx} {ny} 1\n".encode())
            for axis, count in (("X", nx), ("Y", ny), ("Z", 1)):
                out.write(f"{axis}_COORDINATES {count} double\n".encode())
                coords = np.arange(count, dtype=np.float64)
                if self.binary:
                    out.write(coords.astype(">f8").tobytes() + b"\n")
                else:
                    out.write((" ".join("%g" % c for c in coords) +
                               "\n").encode())
            out.write(f"POINT_DATA {n}\n".encode())
            out.write(f"FIELD FieldData {len(data) + 1}\n".encode())
            ids = point_order(np.arange(n, dtype=np.int64))
            _write_array(out, "point_ID", np.ascontiguousarray(ids), 1,
                         self.binary)
            for name, values in data.items():
                components = 1 if values.ndim == 1 else values.shape[1]
                _write_array(out, name,
                             np.ascontiguousarray(point_order(values)),
                             components, self.binary)
        print_log(f"{'VTK conversion done, time':<25}:{time_step:>5}", 0,
                  verbose)
        return path

    def convert_all(self, verbose=True):
        times = _saved_times(self.raw_data_path)
        for t in times:
            self.convert_time(t, verbose=verbose)
        return times


def convert_to_vtk(option, time_step=0, verbose=True, root_dir=".",
                   binary=False):
    """npz2vtk.py:253-290."""
    print_log("-" * 80, 0, verbose)
    print_log("Converting output to VTK...\n", 0, verbose)
    if option not in ("all", "time"):
        raise ValueError("Invalid VTK conversion option!\n"
                         "Supported options: ['all', 'time']")
    op = VTKOperator(_load_metadata(root_dir), root_dir=root_dir,
                     binary=binary, verbose=verbose)
    if option == "all":
        op.convert_all(verbose=verbose)
    else:
        op.convert_time(int(time_step), verbose=verbose)
    print_log("\nConverting output to VTK done!", 0, verbose)
    print_log("-" * 80, 0, verbose)
