#!/usr/bin/env python
"""Builds tuning variants of libplb into pylabolt_b200/lib/variants/ (see
tools/fused_sweep.py); they travel to the GPU box with the tree."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylabolt_b200 import build  # noqa: E402

VARIANTS = {
    # two-steps-per-pass kernel: prefetch ring depth, CTA size, occupancy
    "s0": ["-DPLB_FUSED_STAGES=0"],
    "s2_mb3": ["-DPLB_FUSED_STAGES=2", "-DPLB_FUSED_MINBLOCKS=3"],
    "s3_b64_mb6": ["-DPLB_FUSED_STAGES=3", "-DPLB_FUSED_BLOCK=64", "-DPLB_FUSED_MINBLOCKS=6"],
    "bgk_mb3": ["-DPLB_FUSED_MINBLOCKS_BGK=3"],
}

if __name__ == "__main__":
    names = sys.argv[1:] or sorted(VARIANTS)
    with ThreadPoolExecutor(4) as pool:
        for out in pool.map(lambda n: build.build(force=True, extra_flags=VARIANTS[n],
                                                  variant=n), names):
            print(out)
