#!/usr/bin/env python
"""Builds tuning variants of libplb into pylabolt_b200/lib/variants/ (see
tools/fused_sweep.py); they travel to the GPU box with the tree."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylabolt_b200 import build  # noqa: E402

# Every variant spells out the whole fused-kernel configuration, so that the
# table keeps its meaning when the shipped defaults move (round 2: carry in
# shared memory + one-slot TMA ring + three steps per pass became the default).
def fused(carry, bulk, stages, mb, mb_d3, block=128, extra=()):
    return ["-DPLB_FUSED_CARRY_SMEM=%d" % carry, "-DPLB_FUSED_BULK=%d" % bulk,
            "-DPLB_FUSED_STAGES=%d" % stages, "-DPLB_FUSED_MINBLOCKS=%d" % mb,
            "-DPLB_FUSED_MINBLOCKS_D3=%d" % mb_d3,
            "-DPLB_FUSED_BLOCK=%d" % block] + list(extra)


VARIANTS = {
    # round 1's shipped kernel: carry in registers (162 / 200 registers for two
    # / three steps per pass), cp.async ring of two rows, 3 / 2 CTAs per SM
    "r1": fused(0, 0, 2, 3, 2),
    "r1_s0": fused(0, 0, 0, 3, 2),                 # plain loads, no ring
    "r1_d3_mb3": fused(0, 0, 2, 3, 3),             # depth 3 capped at 168 registers
    "r1_d3_b64_mb5": fused(0, 0, 2, 6, 5, block=64),
    # carry in shared memory, cp.async ring of two rows
    "carry_mb4": fused(1, 0, 2, 4, 3),
    "carry_mb3": fused(1, 0, 2, 3, 3),
    # ring filled by TMA bulk copies (carry in registers): 1 / 2 / 3 / 4 slots
    "bulk_s1": fused(0, 1, 1, 3, 2),
    "bulk_s2": fused(0, 1, 2, 3, 2),
    "bulk_s3": fused(0, 1, 3, 3, 2),
    "bulk_s4": fused(0, 1, 4, 3, 2),
    # carry in shared memory + TMA ring of one slot (18 KB per level and CTA +
    # 18 KB ring): the shipped family.  mbX = CTAs per SM at depth 2, d3mbY at 3
    "cb_s1_mb4": fused(1, 1, 1, 4, 3),             # = the shipped default
    "cb_s1_mb5": fused(1, 1, 1, 5, 4),
    "cb_s1_mb6": fused(1, 1, 1, 6, 4),
    "cb_s1_mb4_d3mb4": fused(1, 1, 1, 4, 4),
    "cb_s1_b64_mb8": fused(1, 1, 1, 8, 6, block=64),
    "cb_s1_b64_mb10": fused(1, 1, 1, 10, 8, block=64),
    "cb_s2_mb4": fused(1, 1, 2, 4, 3),             # two rows in flight, 72 KB at depth 3
    "cb_s2_mb3": fused(1, 1, 2, 3, 3),
    "cb_s1_mb4_late": fused(1, 1, 1, 4, 3, extra=["-DPLB_FUSED_BULK_LATE=1"]),
    "cb_s1_mb4_d3mb2": fused(1, 1, 1, 4, 2),
    "carry_bulk_mb4": fused(1, 1, 2, 4, 3),
    # round 2, second half: the ring filled by ONE rank-3 tensor copy per row
    # (PLB_FUSED_TENSOR), collision constants pinned in registers
    # (PLB_FUSED_PIN), four CTAs of 128 registers at depth 3 -- the shipped
    # default is t1_p1; each variant switches one thing
    "t1_p1": fused(1, 1, 1, 4, 4, extra=["-DPLB_FUSED_TENSOR=1", "-DPLB_FUSED_PIN=1"]),
    "t0_p1": fused(1, 1, 1, 4, 4, extra=["-DPLB_FUSED_TENSOR=0", "-DPLB_FUSED_PIN=1"]),
    "t1_p0": fused(1, 1, 1, 4, 4, extra=["-DPLB_FUSED_TENSOR=1", "-DPLB_FUSED_PIN=0"]),
    "t1_p2": fused(1, 1, 1, 4, 4, extra=["-DPLB_FUSED_TENSOR=1", "-DPLB_FUSED_PIN=2"]),
    "t1_p1_d3mb3": fused(1, 1, 1, 4, 3, extra=["-DPLB_FUSED_TENSOR=1", "-DPLB_FUSED_PIN=1"]),
    "t1_p1_s2": fused(1, 1, 2, 4, 3, extra=["-DPLB_FUSED_TENSOR=1", "-DPLB_FUSED_PIN=1"]),
    "t1_p1_late": fused(1, 1, 1, 4, 4, extra=["-DPLB_FUSED_TENSOR=1", "-DPLB_FUSED_PIN=1",
                                              "-DPLB_FUSED_BULK_LATE=1"]),
    # accuracy of the MRT shortcuts against the oracle over hundreds of steps
    # (bench.py's in-run parity): IEEE division instead of the Newton
    # reciprocal, the two-add pair instead of the three-FMA pair
    "mrt_div": ["-DPLB_MRT_RCP_NEWTON=0"],
    "mrt_pair_old": ["-DPLB_MRT_PAIR_FMA3=0"],
    "mrt_div_pair_old": ["-DPLB_MRT_RCP_NEWTON=0", "-DPLB_MRT_PAIR_FMA3=0"],
    # two ring slots with the per-model CTA counts of the default
    "s2": ["-DPLB_FUSED_STAGES=2"],
    "s2_mrt4": ["-DPLB_FUSED_STAGES=2", "-DPLB_FUSED_MINBLOCKS_D3_MRT=4"],
}

if __name__ == "__main__":
    names = sys.argv[1:] or sorted(VARIANTS)
    with ThreadPoolExecutor(4) as pool:
        for out in pool.map(lambda n: build.build(force=True, extra_flags=VARIANTS[n],
                                                  variant=n), names):
            print(out)
