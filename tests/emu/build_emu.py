"""TEST INFRASTRUCTURE: builds tests/emu/_build/libplb_emu.so.

The libplb sources (pylabolt_b200/csrc/*.cu, unchanged) are compiled by g++
against the stand-in CUDA runtime of tests/emu/include, so that the kernels'
indexing, the warp-shuffle store patterns and the host-side step logic can be
exercised through the same C ABI on a machine without a GPU (-ffp-contract=off:
the arithmetic is that of the -fmad=false build, bit for bit).  This library is
never shipped and never loaded by pylabolt_b200: the product has no CPU path.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, "pylabolt_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
TARGET = os.path.join(OUT_DIR, "libplb_emu.so")
SOURCES = [os.path.join(CSRC, "plb_kernels.cu"), os.path.join(CSRC, "plb_api.cu"),
           os.path.join(HERE, "emu_runtime.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, "plb_internal.h"),
                  os.path.join(CSRC, "plb_collide.cuh"),
                  os.path.join(REPO, "include", "plb.h"),
                  os.path.join(HERE, "include", "cuda_runtime.h"),
                  os.path.join(HERE, "include", "simt.h"),
                  os.path.join(HERE, "include", "nccl.h"),
                  os.path.abspath(__file__)]


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(TARGET):
        t = os.path.getmtime(TARGET)
        if all(os.path.getmtime(d) <= t for d in DEPS):
            return TARGET
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, os.path.basename(src) + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen(
            ["g++", "-x", "c++", "-std=c++17", "-O1", "-g", "-fPIC",
             "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas",
             "-Wno-unused-function", "-I", os.path.join(HERE, "include"),
             "-c", src, "-o", obj]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("emulation build failed")
    subprocess.check_call(["g++", "-shared", "-o", TARGET] + objs + ["-ldl"])
    return TARGET


if __name__ == "__main__":
    print(build(force=True))
