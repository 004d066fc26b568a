#!/usr/bin/env python
"""Times the two-steps-per-pass path of several libplb builds / settings on one
GPU (development tool, not a benchmark).  Each variant is `lib.so` optionally
followed by `:KEY=VALUE,KEY=VALUE` environment overrides (PLB_FUSE,
PLB_FUSED_ROWS, ...):

    python tools/fused_sweep.py [--nx 4096 --ny 16384 --steps 60] \
        pylabolt_b200/lib/libplb.so:PLB_FUSE=0 pylabolt_b200/lib/libplb.so \
        pylabolt_b200/lib/variants/libplb_mb5.so:PLB_FUSED_ROWS=128
"""
import argparse
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CHILD = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, %(repo)r)
from pylabolt_b200 import capi
nx, ny, steps = %(nx)d, %(ny)d, %(steps)d
out = {}
for collision, forcing in %(models)s:
    p = capi.Plb(nx, ny, 1.25, collision=collision, forcing=forcing,
                 gravity=(1e-6, 0.0), x_periodic=True, y_periodic=False)
    size = p.size
    nyp = ny + 2
    bottom = np.arange(1, nx + 1, dtype=np.int64) * nyp + 1
    top = np.arange(1, nx + 1, dtype=np.int64) * nyp + ny
    p.add_boundary_element("bounce_back", bottom, [4, 7, 8], [2, 5, 6], [0, 1])
    p.add_boundary_element("bounce_back", top, [2, 5, 6], [4, 7, 8], [0, -1])
    p.finalize_geometry()
    # a rough density field, so that a misplaced population changes the result
    rho = 1.0 + 1e-3 * ((np.arange(size, dtype=np.int64) * 2654435761 %% 1024) / 1024.0)
    p.upload(capi.DENSITY, rho)
    del rho
    p.initialize_pop()
    p.step(6)
    p.sync()
    p.profile_enable(True)
    p.event_record(0)
    p.step(steps)
    p.event_record(1)
    p.sync()
    ms = p.event_elapsed_ms(0, 1) / steps
    kernel_ms, launches = p.profile_read()
    info = p.fused_info()
    # every variant must leave the same field behind: hash of rho after two
    # more (moment-storing) steps
    p.step(2, store_moments=True)
    import hashlib
    digest = hashlib.blake2b(p.download(capi.DENSITY_INNER).tobytes(),
                             digest_size=6).hexdigest()
    out[f"{collision}/{forcing}"] = {
        "hash": digest,
        "glups": round(nx * ny / (ms * 1e-3) / 1e9, 2),
        "kernel_share": round(kernel_ms / (ms * steps), 3),
        "pairs": info["pairs"], "rows": info["rows"]}
    p.close()
out["build"] = capi.load_library().plb_build_info().decode().split("; ")[1]
print(json.dumps(out))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--ny", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--models", default="mrt,bgk")
    ap.add_argument("--timeout", type=float, default=90.0,
                    help="seconds per variant")
    ap.add_argument("variants", nargs="+")
    args = ap.parse_args()
    models = []
    if "mrt" in args.models:
        models.append(("MRT", "guo_second_order"))
    if "bgk" in args.models:
        models.append(("BGK", None))
    for spec in args.variants:
        lib, _, overrides = spec.partition(":")
        env = dict(os.environ, PLB_LIB=os.path.abspath(lib))
        for kv in filter(None, overrides.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        code = CHILD % {"repo": REPO, "nx": args.nx, "ny": args.ny,
                        "steps": args.steps, "models": repr(models)}
        tag = f"{os.path.basename(lib)} {overrides}"
        try:
            proc = subprocess.run([sys.executable, "-c", code], env=env,
                                  capture_output=True, text=True,
                                  timeout=args.timeout)
        except subprocess.TimeoutExpired:
            # a variant that hangs must not eat the GPU budget of the others
            print(f"{tag:44s} TIMED OUT after {args.timeout} s", flush=True)
            continue
        if proc.returncode != 0:
            print(f"{tag:44s} FAILED {proc.stderr[-300:]}", flush=True)
            continue
        res = json.loads(proc.stdout.strip().splitlines()[-1])
        build = res.pop("build", "")
        cells = "  ".join(f"{k}: {v['glups']:6.2f} GLUPS (kernel share "
                          f"{v['kernel_share']}, pairs {v['pairs']}, rows {v['rows']}, "
                          f"rho hash {v['hash']})"
                          for k, v in res.items())
        print(f"{tag:44s} {cells}  [{build}]", flush=True)


if __name__ == "__main__":
    main()
