#!/usr/bin/env python
"""Builds tuning variants of libplb into pylabolt_b200/lib/variants/ (see
tools/fused_sweep.py); they travel to the GPU box with the tree."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylabolt_b200 import build  # noqa: E402

VARIANTS = {
    # fused kernel (k_bulk_fused): prefetch ring depth, CTA size, occupancy.
    # Depth 3 is selected at run time (PLB_FUSE_DEPTH=3); PLB_FUSED_MINBLOCKS
    # applies to depth 2, depth 3 is capped at two 128-thread CTAs.
    "s0": ["-DPLB_FUSED_STAGES=0"],
    "s2_mb4": ["-DPLB_FUSED_STAGES=2", "-DPLB_FUSED_MINBLOCKS=4"],
    "s3_b64_mb6": ["-DPLB_FUSED_STAGES=3", "-DPLB_FUSED_BLOCK=64", "-DPLB_FUSED_MINBLOCKS=6"],
    "s2_b64_mb6": ["-DPLB_FUSED_STAGES=2", "-DPLB_FUSED_BLOCK=64", "-DPLB_FUSED_MINBLOCKS=6"],
    "s2_b64_mb5": ["-DPLB_FUSED_STAGES=2", "-DPLB_FUSED_BLOCK=64", "-DPLB_FUSED_MINBLOCKS=5"],
    "s3_b64_mb5": ["-DPLB_FUSED_STAGES=3", "-DPLB_FUSED_BLOCK=64", "-DPLB_FUSED_MINBLOCKS=5"],
    # depth 3: 64-thread CTAs, five per SM (204 registers), ring of 2 / 3
    "d3_b64_mb5": ["-DPLB_FUSED_BLOCK=64", "-DPLB_FUSED_MINBLOCKS=6", "-DPLB_FUSED_MINBLOCKS_D3=5"],
    "d3_s3_b64_mb5": ["-DPLB_FUSED_STAGES=3", "-DPLB_FUSED_BLOCK=64", "-DPLB_FUSED_MINBLOCKS=6",
                      "-DPLB_FUSED_MINBLOCKS_D3=5"],
    "d3_mb3": ["-DPLB_FUSED_MINBLOCKS_D3=3"],
    # ring filled by TMA bulk copies on per-warp mbarriers (no LSU instruction,
    # no destination registers): 2 / 3 / 4 slots, and four CTAs per SM for the
    # kernels that then fit 128 registers
    "bulk_s2": ["-DPLB_FUSED_BULK=1"],
    "bulk_s3": ["-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=3"],
    "bulk_s4": ["-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=4"],
    "bulk_s3_mb4": ["-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=3", "-DPLB_FUSED_MINBLOCKS=4"],
    # carried populations in shared memory (18 KB per level and CTA) instead of
    # registers: two steps per pass in 122 registers -> four CTAs per SM (16
    # warps, 55 KB each); three steps per pass in 126 -> three CTAs (12 warps,
    # 74 KB each).  With the TMA ring on top: 94 / 96 registers.
    "carry_mb4": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_MINBLOCKS=4",
                  "-DPLB_FUSED_MINBLOCKS_D3=3"],
    "carry_bulk_mb4": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1",
                       "-DPLB_FUSED_MINBLOCKS=4", "-DPLB_FUSED_MINBLOCKS_D3=3"],
    # 64-thread CTAs: the same warps per SM in twice as many, smaller CTAs
    # (the tail of a wave and the shared-memory granularity are finer)
    "carry_bulk_b64_mb8": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1",
                           "-DPLB_FUSED_BLOCK=64", "-DPLB_FUSED_MINBLOCKS=8",
                           "-DPLB_FUSED_MINBLOCKS_D3=6"],
    "carry_mb3": ["-DPLB_FUSED_CARRY_SMEM=1"],
    # ... and a TMA ring of ONE slot (refilled as soon as it has been read into
    # registers: still one row ahead, 18 KB): 37 KB per CTA at depth 2, 55 KB at
    # depth 3 -> five CTAs per SM (20 warps, 90 registers) / four (16 warps, 94
    # registers); six CTAs (24 warps, the single-step kernel's occupancy) cost
    # 8 bytes of spills at 80 registers
    "cb_s1_mb5": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=1",
                  "-DPLB_FUSED_MINBLOCKS=5", "-DPLB_FUSED_MINBLOCKS_D3=4"],
    "cb_s1_mb6": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=1",
                  "-DPLB_FUSED_MINBLOCKS=6", "-DPLB_FUSED_MINBLOCKS_D3=4"],
    "cb_s1_mb4": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=1",
                  "-DPLB_FUSED_MINBLOCKS=4", "-DPLB_FUSED_MINBLOCKS_D3=3"],
    "cb_s1_b64_mb10": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=1",
                       "-DPLB_FUSED_BLOCK=64", "-DPLB_FUSED_MINBLOCKS=10",
                       "-DPLB_FUSED_MINBLOCKS_D3=8"],
    "bulk_s1": ["-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=1"],
    # round 2: the TMA ring raced on hardware (slot refilled by the async proxy
    # right after the generic-proxy reads).  fence = cross-proxy fence before
    # the refill (now the default), late = refill after the row was collided.
    "cb_s1_mb5_late": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=1",
                       "-DPLB_FUSED_MINBLOCKS=5", "-DPLB_FUSED_MINBLOCKS_D3=4",
                       "-DPLB_FUSED_BULK_LATE=1"],
    "cb_s1_mb5_nofence_late": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1",
                               "-DPLB_FUSED_STAGES=1", "-DPLB_FUSED_MINBLOCKS=5",
                               "-DPLB_FUSED_MINBLOCKS_D3=4", "-DPLB_FUSED_BULK_LATE=1",
                               "-DPLB_FUSED_BULK_FENCE=0"],
    "cb_s1_mb5_d3mb5": ["-DPLB_FUSED_CARRY_SMEM=1", "-DPLB_FUSED_BULK=1", "-DPLB_FUSED_STAGES=1",
                        "-DPLB_FUSED_MINBLOCKS=5", "-DPLB_FUSED_MINBLOCKS_D3=5"],
}

if __name__ == "__main__":
    names = sys.argv[1:] or sorted(VARIANTS)
    with ThreadPoolExecutor(4) as pool:
        for out in pool.map(lambda n: build.build(force=True, extra_flags=VARIANTS[n],
                                                  variant=n), names):
            print(out)
