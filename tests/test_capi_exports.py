"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/plb.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from pylabolt_b200 import build as plb_build
from pylabolt_b200 import capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "plb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plb_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("strict", [False, True])
def test_library_exports_every_declared_symbol(strict):
    plb_build.build()
    path = plb_build.lib_path(strict)
    assert os.path.exists(path), "run python -m pylabolt_b200.build"
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name


def test_python_binding_covers_the_header():
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_config_struct_matches_header_layout():
    """plb_config is plain C: 2 int32, 2 int64, 6 int32, then doubles."""
    assert ctypes.sizeof(capi.PlbConfig) == 8 + 16 + 24 + 8 * (1 + 9 + 2 + 2 + 1 + 9)
    assert capi.PlbConfig.omega.offset == 48
    assert capi.PlbConfig.weights.offset == 48 + 8 * 15


def test_no_cpu_fallback(monkeypatch):
    """A missing library is an error, not a silent fallback."""
    monkeypatch.setattr(plb_build, "LIB_DIR", "/nonexistent")
    monkeypatch.setattr(capi, "_libs", {})
    with pytest.raises(capi.PlbError, match="no CPU fallback"):
        capi.load_library(strict=False)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "pylabolt_b200")
    for root, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, name)).read()
                assert "import oracle" not in text and "from oracle" not in text
                assert "liboracle" not in text


def test_build_info_describes_the_shipped_kernels():
    """plb_build_info needs no device: the shipped libraries are the measured
    configuration (carry in shared memory, one-slot ring filled by one TMA
    tensor copy per row, collision constants pinned, CTAs per SM per depth and
    collision model), production and strict alike."""
    for strict in (False, True):
        text = capi.load_library(strict=strict).plb_build_info().decode()
        assert "fused: block=128 ctas_per_sm=4/4/3/3 (depth2/bgk3/mrt3/mrt4)" in text, text
        assert "stages=1 ring=tma-tensor carry=shared pin=1 smem=dynamic" in text, text
        assert "emulation" not in text


def test_create_fails_loudly_without_a_device():
    """No GPU here: plb_create must return an error code with a message (and
    the Python shim must raise), never hand back a handle that computes
    somewhere else."""
    # (not asked of torch: importing it here would put the real CUDA runtime
    # into this process's global symbol scope)
    if os.path.exists("/dev/nvidiactl"):
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.PlbError, match="libplb error"):
        capi.Plb(16, 16, 1.25)
    lib = capi.load_library(strict=False)
    handle = ctypes.c_void_p()
    cfg = capi.PlbConfig()
    rc = lib.plb_create(ctypes.byref(cfg), ctypes.byref(handle))
    assert rc < 0 and not handle.value
    assert lib.plb_last_error().decode() != ""
