#!/usr/bin/env python
"""Times the step kernels of several libplb builds on one GPU (development
tool, not a benchmark): for each library given on the command line and each
(collision, forcing) pair, GLUPS and the HBM fraction of a periodic channel.

    python tools/kernel_sweep.py [--nx 8192 --ny 16384 --steps 50] lib1.so lib2.so ...
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
for collision, forcing in (("BGK", None), ("BGK", "guo_second_order"),
                           ("MRT", None), ("MRT", "guo_second_order")):
    p = capi.Plb(nx, ny, 1.25, collision=collision, forcing=forcing,
                 gravity=(1e-6, 0.0), x_periodic=True, y_periodic=False)
    size = p.size
    nyp = ny + 2
    bottom = np.arange(1, nx + 1, dtype=np.int64) * nyp + 1
    top = np.arange(1, nx + 1, dtype=np.int64) * nyp + ny
    p.add_boundary_element("bounce_back", bottom, [4, 7, 8], [2, 5, 6], [0, 1])
    p.add_boundary_element("bounce_back", top, [2, 5, 6], [4, 7, 8], [0, -1])
    p.finalize_geometry()
    rho = np.ones(size)
    p.upload(capi.DENSITY, rho)
    del rho
    p.initialize_pop()
    p.step(5)
    p.sync()
    p.event_record(0)
    p.step(steps)
    p.event_record(1)
    p.sync()
    ms = p.event_elapsed_ms(0, 1) / steps
    glups = nx * ny / (ms * 1e-3) / 1e9
    out[f"{collision}/{forcing}"] = round(glups, 2)
    p.close()
print(json.dumps(out))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=8192)
    ap.add_argument("--ny", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("libs", nargs="+")
    args = ap.parse_args()
    peak = 6555.8
    try:
        peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for lib in args.libs:
        env = dict(os.environ, PLB_LIB=os.path.abspath(lib))
        code = CHILD % {"repo": REPO, "nx": args.nx, "ny": args.ny,
                        "steps": args.steps}
        proc = subprocess.run([sys.executable, "-c", code], env=env,
                              capture_output=True, text=True)
        if proc.returncode != 0:
            print(f"{os.path.basename(lib):20s} FAILED {proc.stderr[-300:]}")
            continue
        res = json.loads(proc.stdout.strip().splitlines()[-1])
        cells = "  ".join(f"{k}={v:6.2f} ({v * 144 / peak * 100:5.1f}%)"
                          for k, v in res.items())
        print(f"{os.path.basename(lib):20s} {cells}", flush=True)


if __name__ == "__main__":
    main()
