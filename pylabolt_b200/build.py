"""Builds libplb (the sm_100a CUDA library behind include/plb.h) in-tree.

Two shared objects are produced from the same sources:

  lib/libplb.so         production build; ptxas may contract a*b+c into DFMA
  lib/libplb_strict.so  -fmad=false; bit-identical to the reference's numba
                        kernels on BGK paths, used by the parity tests to tell
                        a rounding difference from a semantic one

    python -m pylabolt_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  The .so files are git-ignored
but travel with the tree to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
REPO = os.path.dirname(HERE)

SOURCES = ["plb_kernels.cu", "plb_api.cu"]
HEADERS = ["plb_internal.h", "plb_collide.cuh",
           os.path.join(REPO, "include", "plb.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "--shared", "-cudart", "static",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libplb cannot be built")
    return nvcc


def lib_path(strict=False):
    return os.path.join(LIB_DIR,
                        "libplb_strict.so" if strict else "libplb.so")


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = [os.path.join(CSRC, s) for s in SOURCES]
    deps += [h if os.path.isabs(h) else os.path.join(CSRC, h)
             for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), variant=None):
    """variant: build lib/variants/libplb_<variant>.so (production flags plus
    extra_flags) instead of the two shipped libraries -- tuning experiments.
    Each library is linked under a temporary name and renamed into place, so
    a tree that is being snapshot never holds a half-written .so; the two
    shipped libraries are compiled side by side."""
    os.makedirs(LIB_DIR, exist_ok=True)
    jobs = []
    for strict in (False, True):
        target = lib_path(strict)
        if variant is not None:
            if strict:
                continue
            os.makedirs(os.path.join(LIB_DIR, "variants"), exist_ok=True)
            target = os.path.join(LIB_DIR, "variants", f"libplb_{variant}.so")
        if not force and not _stale(target):
            continue
        tmp = target + f".tmp{os.getpid()}"
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags)
        if strict:
            cmd += ["-fmad=false"]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += [os.path.join(CSRC, s) for s in SOURCES]
        cmd += ["-o", tmp, "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        jobs.append((target, tmp, subprocess.Popen(cmd, cwd=CSRC)))
    built = []
    failed = None
    for target, tmp, proc in jobs:
        if proc.wait() != 0:
            failed = failed or target
            if os.path.exists(tmp):
                os.unlink(tmp)
            continue
        os.replace(tmp, target)
        built.append(target)
    if failed:
        raise subprocess.CalledProcessError(1, "nvcc (" + failed + ")")
    return built


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    for p in out:
        print("built", os.path.relpath(p, REPO))
    if not out:
        print("libplb is up to date")
