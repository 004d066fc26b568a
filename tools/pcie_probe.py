#!/usr/bin/env python
"""Host <-> device transfer rates on this box (development tool): torch pinned
copies as the ceiling, then plb_upload / plb_download of rho and u."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from pylabolt_b200 import capi

n = 8192 * 16384 * 3
host = torch.empty(n, dtype=torch.float64).pin_memory()
dev = torch.empty(n, dtype=torch.float64, device="cuda")
for name, dst, src in (("torch H2D", dev, host), ("torch D2H", host, dev)):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{name}: {n * 8 / dt / 1e9:.1f} GB/s", flush=True)
# chunked D2H like libplb (64 MB pieces on one stream)
chunk = (64 << 20) // 8
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for o in range(0, n, chunk):
        host[o:o + chunk].copy_(dev[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"torch D2H 64MB chunks: {n * 8 / dt / 1e9:.1f} GB/s", flush=True)
del host, dev
torch.cuda.empty_cache()

p = capi.Plb(8192, 16384, 1.25, x_periodic=True)
size = p.size
bufs = {"rho": p.pinned((size,)), "u": p.pinned((size, 2)),
        "rho_in": p.pinned((p.nx * p.ny,)), "u_in": p.pinned((p.nx * p.ny, 2))}
for b in bufs.values():
    b.array[:] = 1.0
for rep in range(2):
    for name, fn, field, buf in (
            ("upload rho", p.upload, capi.DENSITY, bufs["rho"]),
            ("upload u", p.upload, capi.VELOCITY, bufs["u"]),
            ("download rho padded", p.download, capi.DENSITY, bufs["rho"]),
            ("download u padded", p.download, capi.VELOCITY, bufs["u"]),
            ("download rho inner", p.download, capi.DENSITY_INNER, bufs["rho_in"]),
            ("download u inner", p.download, capi.VELOCITY_INNER, bufs["u_in"])):
        p.sync(); t0 = time.perf_counter()
        fn(field, buf.array); p.sync()
        dt = time.perf_counter() - t0
        print(f"{name}: {buf.array.nbytes / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms)", flush=True)
