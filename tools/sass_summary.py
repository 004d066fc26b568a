#!/usr/bin/env python
"""Static evidence for the hot kernels of libplb.so: registers / shared memory /
spills (cuobjdump --dump-resource-usage) and the SASS instruction mix of the
kernel bodies (cuobjdump -sass), without a GPU.

    python tools/sass_summary.py [lib.so] > profiles/<round>_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    REPO, "pylabolt_b200", "lib", "libplb.so")
# (collision, forcing): 0/0 = BGK without forcing (cavity), 2/2 = two-stress-
# moment MRT + Guo second order (channel)
KERNELS = [r"k_bulk_fused<2, 2, 4>", r"k_bulk_fused<0, 0, 3>", r"k_bulk_fused<2, 2, 3>",
           r"k_bulk_fused<2, 2, 2>",
           r"k_bulk_vec2<2, 2, false>", r"k_bulk_vec2<0, 0, false>",
           r"k_links<2, 2, false>", r"k_face_unpack", r"k_face_signal"]
GROUPS = [
    ("fp64 arithmetic", r"^(DFMA|DMUL|DADD|DSETP|MUFU\.RCP64H)"),
    ("global loads (LDG)", r"^LDG"),
    ("  of which 128-bit", r"^LDG\.E\.128|^LDG.*\.128"),
    ("async global->shared (LDGSTS)", r"^LDGSTS"),
    ("TMA bulk copies (UBLKCP)", r"^UBLKCP"),
    ("TMA tensor copies (UTMALDG)", r"^UTMALDG"),
    ("shared loads (LDS)", r"^LDS"),
    ("global stores (STG)", r"^STG"),
    ("  of which 128-bit", r"^STG.*\.128"),
    ("local memory (LDL/STL = spills)", r"^(LDL|STL)"),
    ("warp shuffles (SHFL)", r"^SHFL"),
    ("votes (VOTE)", r"^VOTE"),
    ("barriers (BAR / SYNCS)", r"^(BAR|SYNCS)"),
    ("atomics (ATOM/RED)", r"^(ATOM|RED)"),
]


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def resources():
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB],
                         capture_output=True, text=True).stdout
    res = {}
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = demangle(m.group(1))
            continue
        if name and "REG:" in line:
            res[name] = line.strip()
            name = None
    return res


def sass():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    bodies = {}
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = demangle(m.group(1))
            bodies[name] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            bodies[name].append(m.group(1))
    return bodies


def main():
    res, bodies = resources(), sass()
    print("# SASS summary of", os.path.relpath(LIB, REPO), "(sm_100a, static; no GPU needed)")
    for pattern in KERNELS:
        for name in sorted(bodies):
            if not re.search(pattern, name):
                continue
            ops = bodies[name]
            print("\n##", name.split("(")[0])
            print("  ", res.get(name, "(no resource line)"))
            print(f"   {len(ops)} instructions")
            for label, rx in GROUPS:
                n = sum(1 for op in ops if re.match(rx, op))
                if n:
                    print(f"   {n:6d}  {label}")
            top = collections.Counter(op.split(".")[0] for op in ops).most_common(8)
            print("   top opcodes:", ", ".join(f"{k} {v}" for k, v in top))


if __name__ == "__main__":
    main()
