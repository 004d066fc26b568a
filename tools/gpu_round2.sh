#!/bin/bash
# First GPU round trip of the NEXT round (run under gpurun, ~6 min on one GPU):
# what round 1 could not measure any more.  Build the variants first:
#   python tools/build_variants.py        (-> pylabolt_b200/lib/variants/)
# Outputs land in gpurun_out/<tag>_*; copy what is to be judged to profiles/.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
L=pylabolt_b200/lib
V=$L/variants
el start
# 1. parity of everything that changed after the last GPU run of round 1
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_zz_fused_depth3.py -x -q 2>&1 | tail -5 > $out/${tag}_pytest_fused.log
el "pytest fused: $(tail -1 $out/${tag}_pytest_fused.log)"
# 2. two vs three steps per pass, occupancy / ring variants (each line ~7 s)
timeout 160 python tools/fused_sweep.py $L/libplb.so:PLB_FUSE=0 $L/libplb.so $L/libplb.so:PLB_FUSE_DEPTH=3 > $out/${tag}_sweep.txt 2>&1
# nine-rate MRT (the literal 9x9 transform), single-step and fused
timeout 120 python tools/fused_sweep.py --models mrt $L/libplb.so:PLB_MRT_GENERAL=1,PLB_FUSE=0 $L/libplb.so:PLB_MRT_GENERAL=1 \
    $L/libplb.so:PLB_MRT_GENERAL=1,PLB_FUSE_DEPTH=3 $V/libplb_cb_s1_mb5.so:PLB_MRT_GENERAL=1 $V/libplb_cb_s1_mb4.so:PLB_MRT_GENERAL=1 \
    $V/libplb_carry_mb4.so:PLB_MRT_GENERAL=1 >> $out/${tag}_sweep.txt 2>&1
timeout 800 python tools/fused_sweep.py --models mrt \
    $V/libplb_cb_s1_mb5.so $V/libplb_cb_s1_mb6.so $V/libplb_cb_s1_mb4.so $V/libplb_cb_s1_b64_mb10.so \
    $V/libplb_cb_s1_mb5.so:PLB_FUSE_DEPTH=3 $V/libplb_cb_s1_mb5.so:PLB_FUSE_DEPTH=3,PLB_FUSED_ROWS=64 \
    $V/libplb_cb_s1_b64_mb10.so:PLB_FUSE_DEPTH=3 $V/libplb_cb_s1_mb5.so:PLB_FUSED_ROWS=64 \
    $V/libplb_cb_s1_mb5.so:PLB_FUSED_DYNAMIC=1,PLB_FUSED_ROWS=128 $V/libplb_bulk_s1.so \
    $L/libplb.so:PLB_FUSE_DEPTH=3,PLB_FUSED_ROWS=16 $L/libplb.so:PLB_FUSE_DEPTH=3,PLB_FUSED_ROWS=64 \
    $V/libplb_d3_b64_mb5.so:PLB_FUSE_DEPTH=3 $V/libplb_d3_s3_b64_mb5.so:PLB_FUSE_DEPTH=3 \
    $V/libplb_d3_mb3.so:PLB_FUSE_DEPTH=3 \
    $V/libplb_s3_b64_mb6.so $V/libplb_s2_b64_mb6.so $V/libplb_s2_b64_mb5.so $V/libplb_s3_b64_mb5.so \
    $V/libplb_s2_mb4.so $V/libplb_s0.so \
    $V/libplb_bulk_s2.so $V/libplb_bulk_s3.so $V/libplb_bulk_s4.so $V/libplb_bulk_s3_mb4.so \
    $V/libplb_bulk_s3.so:PLB_FUSE_DEPTH=3 $V/libplb_bulk_s3.so:PLB_FUSED_DYNAMIC=1,PLB_FUSED_ROWS=128 \
    $V/libplb_carry_mb4.so $V/libplb_carry_bulk_mb4.so $V/libplb_carry_bulk_b64_mb8.so $V/libplb_carry_mb3.so \
    $V/libplb_carry_mb4.so:PLB_FUSE_DEPTH=3 $V/libplb_carry_bulk_mb4.so:PLB_FUSE_DEPTH=3 \
    $V/libplb_carry_bulk_b64_mb8.so:PLB_FUSE_DEPTH=3 $V/libplb_carry_bulk_mb4.so:PLB_FUSE_DEPTH=3,PLB_FUSED_ROWS=64 \
    $V/libplb_carry_bulk_mb4.so:PLB_FUSED_DYNAMIC=1,PLB_FUSED_ROWS=128 \
    $L/libplb.so:PLB_FUSED_DYNAMIC=1 $L/libplb.so:PLB_FUSED_DYNAMIC=1,PLB_FUSED_ROWS=64 \
    $L/libplb.so:PLB_FUSED_DYNAMIC=1,PLB_FUSED_ROWS=128 $L/libplb.so:PLB_FUSED_DYNAMIC=1,PLB_FUSED_ROWS=256 \
    $L/libplb.so:PLB_FUSED_ROWS=64 $L/libplb.so:PLB_FUSED_ROWS=128 \
    $L/libplb.so:PLB_FUSE_DEPTH=3,PLB_FUSED_DYNAMIC=1,PLB_FUSED_ROWS=128 >> $out/${tag}_sweep.txt 2>&1
el sweep
# 3. the headline line, default and depth 3
timeout 500 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
el bench
PLB_FUSE_DEPTH=3 timeout 200 python bench.py --steps 102 --no-extras --no-cpu-baseline \
    > $out/${tag}_bench_depth3.json 2> $out/${tag}_bench_depth3.err
el bench-depth3
# 4. ncu: the shipped two-step build (its round-1 traffic figure is another build's) and depth 3
for d in 2 3; do
  PLB_FUSE_DEPTH=$d timeout 150 ncu --set full --clock-control none --import-source on \
      -k regex:k_bulk_fused -s 3 -c 1 -f -o $out/${tag}_ncu_channel_depth$d \
      python bench.py --steps 6 --warmup 6 --no-extras --no-cpu-baseline --no-parity > $out/${tag}_ncu_bench_$d.log 2>&1
  ncu -i $out/${tag}_ncu_channel_depth$d.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_channel_depth$d.csv 2>/dev/null
  ncu -i $out/${tag}_ncu_channel_depth$d.ncu-rep --page details > $out/${tag}_ncu_details_channel_depth$d.txt 2>/dev/null
done
el ncu
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_channel.csv \
    python bench.py --steps 4 --warmup 4 --no-extras --no-cpu-baseline --no-parity > /dev/null 2>&1
el launch-list
# 5. the rest of the GPU suite
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest_all.log
el "pytest all: $(tail -1 $out/${tag}_pytest_all.log)"
cat $out/${tag}_timeline.log $out/${tag}_sweep.txt
head -c 400 $out/${tag}_bench.json
