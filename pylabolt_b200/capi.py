"""ctypes binding of libplb (include/plb.h).

There is no CPU fallback: if the CUDA library has not been built, or no CUDA
device is visible, every entry point raises.  Build with
``python -m pylabolt_b200.build``.
"""
import ctypes
import os

import numpy as np

from . import build as _build

PLB_ABI_VERSION = 1

# enum plb_field
SOLID, DENSITY, VELOCITY, POP, DENSITY_INNER, VELOCITY_INNER = range(6)
# enum plb_collision / plb_forcing / plb_bc_type
COLLISION = {"BGK": 0, "MRT": 1}
FORCING = {None: 0, "guo_linear": 1, "guo_second_order": 2}
BC_TYPE = {"bounce_back": 0, "fixed_velocity": 1, "fixed_pressure": 2,
           "periodic": 3, "zero_gradient": 4}

EXPORTS = [
    "plb_create", "plb_destroy", "plb_last_error",
    "plb_add_boundary_element", "plb_finalize_geometry",
    "plb_upload", "plb_download", "plb_fill", "plb_initialize_pop",
    "plb_step", "plb_sync", "plb_residue_sums",
    "plb_comm_unique_id", "plb_comm_init",
    "plb_event_record", "plb_event_elapsed_ms", "plb_kernel_launches",
    "plb_host_alloc", "plb_host_free", "plb_flush_l2",
    "plb_profile_enable", "plb_profile_read", "plb_info",
    "plb_link_nodes", "plb_download_link_exchange", "plb_copy_bandwidth",
    "plb_device_pci_bus_id", "plb_fused_info", "plb_memory_info", "plb_build_info",
]
STORE_MOMENTS, RECORD_LINKS = 1, 2
# plb_info()["faces"]: how the slab-face populations travel
FACES_NONE, FACES_SELF, FACES_NCCL, FACES_P2P = range(4)


class PlbConfig(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("device", ctypes.c_int32),
        ("nx", ctypes.c_int64), ("ny", ctypes.c_int64),
        ("x_periodic", ctypes.c_int32), ("y_periodic", ctypes.c_int32),
        ("left_neighbor", ctypes.c_int32), ("right_neighbor", ctypes.c_int32),
        ("collision", ctypes.c_int32), ("forcing", ctypes.c_int32),
        ("omega", ctypes.c_double),
        ("mrt_rates", ctypes.c_double * 9),
        ("gravity", ctypes.c_double * 2),
        ("inv_cs_2", ctypes.c_double), ("inv_cs_4", ctypes.c_double),
        ("float_min", ctypes.c_double),
        ("weights", ctypes.c_double * 9),
    ]


class PlbError(RuntimeError):
    pass


_libs = {}

# The product only ever runs the sm_100a build.  tests/emu/ holds a host
# emulation of the same sources (kernel-logic tests without a GPU); its test
# harness -- tests/conftest.py, never the product -- flips this switch.  With
# it off, a PLB_LIB that points at an emulated build is refused.
_accept_emulated_build = False


def load_library(strict=None):
    """dlopen libplb.so (or libplb_strict.so when strict / PLB_STRICT=1).
    PLB_LIB selects another sm_100a build of libplb (tuning variants)."""
    if strict is None:
        strict = os.environ.get("PLB_STRICT", "0") not in ("", "0")
    strict = bool(strict)
    path = os.environ.get("PLB_LIB") or _build.lib_path(strict)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise PlbError(
            f"{path} is missing: the b200 back end has no CPU fallback. "
            "Build it with `python -m pylabolt_b200.build`.")
    lib = ctypes.CDLL(path)
    vp, i32, i64, dbl = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64,
                         ctypes.c_double)
    lib.plb_last_error.restype = ctypes.c_char_p
    lib.plb_create.argtypes = [ctypes.POINTER(PlbConfig), ctypes.POINTER(vp)]
    lib.plb_destroy.argtypes = [vp]
    lib.plb_destroy.restype = None
    lib.plb_add_boundary_element.argtypes = [
        vp, i32, ctypes.POINTER(i64), i64, ctypes.POINTER(i64),
        ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(dbl), dbl]
    lib.plb_finalize_geometry.argtypes = [vp]
    lib.plb_upload.argtypes = [vp, i32, vp, ctypes.c_size_t]
    lib.plb_download.argtypes = [vp, i32, vp, ctypes.c_size_t]
    lib.plb_fill.argtypes = [vp, i32, ctypes.POINTER(dbl)]
    lib.plb_initialize_pop.argtypes = [vp]
    lib.plb_step.argtypes = [vp, i64, i32]
    lib.plb_sync.argtypes = [vp]
    lib.plb_residue_sums.argtypes = [vp, ctypes.POINTER(dbl)]
    lib.plb_comm_unique_id.argtypes = [vp]
    lib.plb_comm_init.argtypes = [vp, vp, i32, i32, i32, i32]
    lib.plb_event_record.argtypes = [vp, i32]
    lib.plb_event_elapsed_ms.argtypes = [vp, i32, i32,
                                         ctypes.POINTER(ctypes.c_float)]
    lib.plb_kernel_launches.argtypes = [vp, i32]
    lib.plb_kernel_launches.restype = i64
    lib.plb_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    lib.plb_host_free.argtypes = [vp]
    lib.plb_flush_l2.argtypes = [vp]
    lib.plb_profile_enable.argtypes = [vp, i32]
    lib.plb_profile_read.argtypes = [vp, ctypes.POINTER(dbl),
                                     ctypes.POINTER(i64)]
    lib.plb_info.argtypes = [vp, ctypes.POINTER(i64)]
    lib.plb_fused_info.argtypes = [vp, ctypes.POINTER(i64)]
    lib.plb_memory_info.argtypes = [vp, ctypes.POINTER(i64)]
    lib.plb_link_nodes.argtypes = [vp, ctypes.POINTER(i64), i64,
                                   ctypes.POINTER(i64)]
    lib.plb_download_link_exchange.argtypes = [vp, ctypes.POINTER(dbl), i64]
    lib.plb_copy_bandwidth.argtypes = [vp, ctypes.POINTER(dbl)]
    lib.plb_build_info.argtypes = []
    lib.plb_build_info.restype = ctypes.c_char_p
    lib.plb_device_pci_bus_id.argtypes = [i32, ctypes.c_char_p, i32]
    info = (lib.plb_build_info() or b"").decode()
    if "emulation" in info and not _accept_emulated_build:
        raise PlbError(
            f"{path} is a host emulation of libplb (test infrastructure): "
            "the b200 back end has no CPU path and will not load it.")
    _libs[path] = lib
    return lib


def device_pci_bus_id(device, strict=None):
    """'0000:1b:00.0' of a CUDA device ordinal."""
    lib = load_library(strict)
    buf = ctypes.create_string_buffer(64)
    if lib.plb_device_pci_bus_id(int(device), buf, 64) != 0:
        raise PlbError(lib.plb_last_error().decode())
    return buf.value.decode().lower()


def _i64(values):
    arr = (ctypes.c_int64 * len(values))()
    arr[:] = [int(v) for v in values]
    return arr


class PinnedArray:
    """numpy view of page-locked host memory from plb_host_alloc."""

    def __init__(self, lib, shape, dtype):
        self._lib = lib
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        self.nbytes = max(1, n * self.dtype.itemsize)
        ptr = ctypes.c_void_p()
        rc = lib.plb_host_alloc(ctypes.byref(ptr), self.nbytes)
        if rc != 0:
            raise PlbError(lib.plb_last_error().decode())
        self._ptr = ptr
        buf = (ctypes.c_char * self.nbytes).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=n).reshape(shape)

    def free(self):
        if self._ptr is not None:
            self.array = None
            self._lib.plb_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Plb:
    """One rank's libplb solver handle."""

    def __init__(self, nx, ny, omega, *, device=0, collision="BGK",
                 forcing=None, gravity=(0.0, 0.0), x_periodic=False,
                 y_periodic=False, left_neighbor=None, right_neighbor=None,
                 mrt_rates=None, lattice=None, float_min=None, strict=None):
        self.lib = load_library(strict)
        if lattice is None:
            from .state import d2q9_constants
            lattice = d2q9_constants()
        cfg = PlbConfig()
        cfg.abi_version = PLB_ABI_VERSION
        cfg.device = int(device)
        cfg.nx, cfg.ny = int(nx), int(ny)
        cfg.x_periodic, cfg.y_periodic = int(bool(x_periodic)), int(bool(y_periodic))
        cfg.left_neighbor = int(bool(x_periodic if left_neighbor is None
                                     else left_neighbor))
        cfg.right_neighbor = int(bool(x_periodic if right_neighbor is None
                                      else right_neighbor))
        cfg.collision = COLLISION[collision]
        cfg.forcing = FORCING[forcing]
        cfg.omega = float(omega)
        if mrt_rates is None:
            mrt_rates = [1.0] * 7 + [float(omega)] * 2
        cfg.mrt_rates[:] = [float(v) for v in mrt_rates]
        cfg.gravity[:] = [float(gravity[0]), float(gravity[1])]
        cfg.inv_cs_2 = float(lattice["inv_cs_2"])
        cfg.inv_cs_4 = float(lattice["inv_cs_4"])
        cfg.float_min = float(np.finfo(np.float64).eps if float_min is None
                              else float_min)
        cfg.weights[:] = [float(v) for v in lattice["weights"]]
        self.config = cfg
        self.nx, self.ny = cfg.nx, cfg.ny
        self.shape = (self.nx + 2, self.ny + 2)
        self.size = self.shape[0] * self.shape[1]
        self._h = ctypes.c_void_p()
        self._check(self.lib.plb_create(ctypes.byref(cfg), ctypes.byref(self._h)))

    # -- plumbing --------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise PlbError(f"libplb error {rc}: "
                           f"{self.lib.plb_last_error().decode()}")

    def close(self):
        if self._h:
            self.lib.plb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- geometry ----------------------------------------------------------
    def add_boundary_element(self, bc_type, boundary_nodes, out_list, inv_list,
                             normal, vector=(0.0, 0.0), scalar=0.0):
        nodes = np.ascontiguousarray(boundary_nodes, dtype=np.int64)
        vec = (ctypes.c_double * 2)(float(vector[0]), float(vector[1]))
        code = BC_TYPE[bc_type] if isinstance(bc_type, str) else int(bc_type)
        self._check(self.lib.plb_add_boundary_element(
            self._h, code,
            nodes.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
            nodes.shape[0], _i64(out_list), _i64(inv_list), _i64(normal), vec,
            float(scalar)))

    def finalize_geometry(self):
        self._check(self.lib.plb_finalize_geometry(self._h))

    # -- data --------------------------------------------------------------
    def upload(self, field, array):
        dtype = np.uint8 if field == SOLID else np.float64
        a = np.ascontiguousarray(array, dtype=dtype)
        self._check(self.lib.plb_upload(self._h, field, a.ctypes.data,
                                        a.nbytes))

    def fill(self, field, value):
        """Uniform DENSITY (scalar) / VELOCITY (pair) on the interior nodes,
        produced on the device (plb_fill)."""
        vals = np.atleast_1d(np.asarray(value, dtype=np.float64))
        buf = (ctypes.c_double * 2)(*(list(vals) + [0.0])[:2])
        self._check(self.lib.plb_fill(self._h, field, buf))

    def download(self, field, out=None):
        shapes = {SOLID: ((self.size,), np.uint8),
                  DENSITY: ((self.size,), np.float64),
                  VELOCITY: ((self.size, 2), np.float64),
                  POP: ((self.size, 9), np.float64),
                  DENSITY_INNER: ((self.nx * self.ny,), np.float64),
                  VELOCITY_INNER: ((self.nx * self.ny, 2), np.float64)}
        shape, dtype = shapes[field]
        if out is None:
            out = np.empty(shape, dtype=dtype)
        assert out.flags.c_contiguous and out.dtype == dtype
        self._check(self.lib.plb_download(self._h, field, out.ctypes.data,
                                          out.nbytes))
        return out

    def initialize_pop(self):
        self._check(self.lib.plb_initialize_pop(self._h))

    # -- the hot path --------------------------------------------------------
    def step(self, n_steps=1, store_moments=False, record_links=False):
        flags = (STORE_MOMENTS if store_moments else 0) | \
            (RECORD_LINKS if record_links else 0)
        self._check(self.lib.plb_step(self._h, int(n_steps), flags))

    def link_nodes(self):
        """Padded flat indices of the link nodes, in list order."""
        n = ctypes.c_int64()
        self._check(self.lib.plb_link_nodes(self._h, None, 0, ctypes.byref(n)))
        out = np.empty(n.value, dtype=np.int64)
        if n.value:
            self._check(self.lib.plb_link_nodes(
                self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                n.value, ctypes.byref(n)))
        return out

    def link_exchange(self, n_links):
        """(n_links, 8): pop[k] + pop_new[inv k] per link node and k = 1..8."""
        out = np.zeros((n_links, 8), dtype=np.float64)
        if n_links:
            self._check(self.lib.plb_download_link_exchange(
                self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                out.size))
        return out

    def sync(self):
        self._check(self.lib.plb_sync(self._h))

    def residue_sums(self):
        out = (ctypes.c_double * 6)()
        self._check(self.lib.plb_residue_sums(self._h, out))
        return np.array(out[:], dtype=np.float64)

    # -- multi GPU -------------------------------------------------------------
    def comm_unique_id(self):
        buf = (ctypes.c_char * 128)()
        self._check(self.lib.plb_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id, rank, n_ranks, left_rank, right_rank):
        buf = (ctypes.c_char * 128).from_buffer_copy(unique_id)
        self._check(self.lib.plb_comm_init(self._h, buf, rank, n_ranks,
                                           left_rank, right_rank))

    # -- measurement -------------------------------------------------------------
    def event_record(self, slot):
        self._check(self.lib.plb_event_record(self._h, slot))

    def event_elapsed_ms(self, start, stop):
        ms = ctypes.c_float()
        self._check(self.lib.plb_event_elapsed_ms(self._h, start, stop,
                                                  ctypes.byref(ms)))
        return float(ms.value)

    def kernel_launches(self, reset=False):
        return int(self.lib.plb_kernel_launches(self._h, int(reset)))

    def profile_enable(self, enable=True):
        self._check(self.lib.plb_profile_enable(self._h, int(bool(enable))))

    def profile_read(self):
        """(summed bulk-kernel ms, launches) since the last read."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        self._check(self.lib.plb_profile_read(self._h, ctypes.byref(ms),
                                              ctypes.byref(n)))
        return float(ms.value), int(n.value)

    def info(self):
        out = (ctypes.c_int64 * 8)()
        self._check(self.lib.plb_info(self._h, out))
        keys = ("n_bulk", "n_link", "n_solid", "pitch", "plane", "variant",
                "n_bulk_timed", "faces")
        return dict(zip(keys, out[:8]))

    def fused_info(self):
        """State of the several-steps-per-pass path (plb_fused_info)."""
        out = (ctypes.c_int64 * 10)()
        self._check(self.lib.plb_fused_info(self._h, out))
        keys = ("active", "n_deep", "n_deep3", "n_list1", "pairs", "rows",
                "strips", "triples", "n_deep4", "quads")
        return dict(zip(keys, out[:10]))

    def memory_info(self):
        """Device bytes held by this handle (plb_memory_info)."""
        out = (ctypes.c_int64 * 6)()
        self._check(self.lib.plb_memory_info(self._h, out))
        keys = ("lattices", "moments", "scratch", "flags_lists_staging",
                "device_free", "device_total")
        return dict(zip(keys, out[:6]))

    def build_info(self):
        """Compile-time kernel configuration of the loaded library."""
        return self.lib.plb_build_info().decode()

    def copy_bandwidth(self):
        """GB/s (read + write) of a device-to-device copy on this GPU, now."""
        gbs = ctypes.c_double()
        self._check(self.lib.plb_copy_bandwidth(self._h, ctypes.byref(gbs)))
        return float(gbs.value)

    def flush_l2(self):
        self._check(self.lib.plb_flush_l2(self._h))

    def pinned(self, shape, dtype=np.float64):
        return PinnedArray(self.lib, shape, dtype)
