#!/bin/bash
# Variant sweep (tools/build_variants.py first): gpurun -- bash tools/gpu_sweep_variants.sh [tag]
# Every line must show the same rho hash per model; TMA variants run twice (a
# race shows as run-to-run different hashes).
tag=${1:-r2c}
out=gpurun_out
mkdir -p $out
L=pylabolt_b200/lib
V=$L/variants
D3=PLB_FUSE_DEPTH=3,PLB_FUSED_ROWS=64
timeout 1300 python tools/fused_sweep.py --models mrt,bgk \
    $L/libplb.so $L/libplb.so:$D3 \
    $V/libplb_carry_mb4.so:$D3 $V/libplb_carry_mb4.so:PLB_FUSE_DEPTH=3,PLB_FUSED_ROWS=128 \
    $V/libplb_d3_mb3.so:$D3 $V/libplb_d3_b64_mb5.so:$D3 \
    $V/libplb_cb_s1_mb5.so:$D3 $V/libplb_cb_s1_mb5.so:$D3 \
    $V/libplb_cb_s1_mb5_late.so:$D3 $V/libplb_cb_s1_mb5_late.so:$D3 \
    $V/libplb_cb_s1_mb5_nofence_late.so:$D3 $V/libplb_cb_s1_mb5_nofence_late.so:$D3 \
    $V/libplb_cb_s1_mb5_d3mb5.so:$D3 \
    $V/libplb_cb_s1_mb4.so:$D3 \
    $V/libplb_cb_s1_mb5.so $V/libplb_cb_s1_mb5_late.so \
    $V/libplb_bulk_s1.so $V/libplb_bulk_s2.so $V/libplb_bulk_s4.so $V/libplb_bulk_s4.so \
    $V/libplb_carry_bulk_mb4.so:$D3 \
    > $out/${tag}_sweep.txt 2>&1
timeout 200 python tools/fused_sweep.py --models mrt \
    $L/libplb.so:PLB_MRT_GENERAL=1,PLB_FUSE=0 $L/libplb.so:PLB_MRT_GENERAL=1 \
    $L/libplb.so:PLB_MRT_GENERAL=1,$D3 $V/libplb_carry_mb4.so:PLB_MRT_GENERAL=1,$D3 \
    $V/libplb_carry_mb4.so:PLB_MRT_GENERAL=1 >> $out/${tag}_sweep.txt 2>&1
cut -c1-260 $out/${tag}_sweep.txt
