#!/bin/bash
# Single-GPU round trip after the tensor-TMA / pinned-constants / predicated-store
# rework of k_bulk_fused: the whole -m gpu suite on the new default, the variant
# sweep that isolates each change, bench.py, ncu of the new default on both
# headline workloads, launch list.
#   python tools/build_variants.py t0_p1 t1_p0 t1_p2 t1_p1_d3mb3 t1_p1_s2 t1_p1_late
#   gpurun --timeout 1800 -- bash tools/gpu_round2e.sh [tag]
tag=${1:-r2e}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
L=pylabolt_b200/lib
V=$L/variants
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/${tag}_pytest_all.log
el "pytest all: $(tail -1 $out/${tag}_pytest_all.log)"
timeout 600 python tools/fused_sweep.py --models mrt,bgk \
    $L/libplb.so $L/libplb.so \
    $V/libplb_t0_p1.so $V/libplb_t1_p0.so $V/libplb_t1_p2.so $V/libplb_t1_p1_d3mb3.so \
    $V/libplb_t1_p1_s2.so $V/libplb_t1_p1_late.so \
    $L/libplb.so:PLB_FUSED_ROWS=32 $L/libplb.so:PLB_FUSED_ROWS=96 $L/libplb.so:PLB_FUSED_ROWS=128 \
    $L/libplb.so:PLB_FUSED_DYNAMIC=1,PLB_FUSED_ROWS=128 \
    $L/libplb.so:PLB_TMA_L2_PROMOTION=0 $L/libplb.so:PLB_TMA_L2_PROMOTION=3 \
    $L/libplb.so:PLB_FUSE_DEPTH=2 $L/libplb.so:PLB_FUSE=0 \
    $L/libplb.so:PLB_MRT_GENERAL=1 \
    > $out/${tag}_sweep.txt 2>&1
el sweep
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
el bench
for wl in channel cavity; do
  timeout 240 ncu --set full --clock-control none --import-source on \
      -k regex:k_bulk_fused -s 3 -c 1 -f -o $out/${tag}_ncu_${wl}_fused3 \
      python bench.py --workload $wl --steps 6 --warmup 6 --no-extras --no-cpu-baseline --no-parity > $out/${tag}_ncu_bench_$wl.log 2>&1
  ncu -i $out/${tag}_ncu_${wl}_fused3.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_${wl}_fused3.csv 2>/dev/null
  ncu -i $out/${tag}_ncu_${wl}_fused3.ncu-rep --page details > $out/${tag}_ncu_details_${wl}_fused3.txt 2>/dev/null
done
el ncu
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $out/${tag}_launches_channel.csv \
    python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline --no-parity > /dev/null 2>&1
el launch-list
cat $out/${tag}_timeline.log
tail -3 $out/${tag}_pytest_all.log
cut -c1-220 $out/${tag}_sweep.txt
head -c 400 $out/${tag}_bench.json
