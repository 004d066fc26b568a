#!/usr/bin/env python
"""Builds tuning variants of libplb into pylabolt_b200/lib/variants/ (see
tools/fused_sweep.py); they travel to the GPU box with the tree."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylabolt_b200 import build  # noqa: E402

VARIANTS = {
    "mb3": ["-DPLB_FUSED_MINBLOCKS=3"],
    "mb5": ["-DPLB_FUSED_MINBLOCKS=5"],
    "mb4_pf2": ["-DPLB_FUSED_L2_AHEAD=2"],
    "mb4_pf4": ["-DPLB_FUSED_L2_AHEAD=4"],
    "mb3_pf2": ["-DPLB_FUSED_MINBLOCKS=3", "-DPLB_FUSED_L2_AHEAD=2"],
    "blk64": ["-DPLB_FUSED_BLOCK=64", "-DPLB_FUSED_MINBLOCKS=8"],
    "blk256": ["-DPLB_FUSED_BLOCK=256", "-DPLB_FUSED_MINBLOCKS=2"],
}

if __name__ == "__main__":
    names = sys.argv[1:] or sorted(VARIANTS)
    with ThreadPoolExecutor(4) as pool:
        for out in pool.map(lambda n: build.build(force=True, extra_flags=VARIANTS[n],
                                                  variant=n), names):
            print(out)
