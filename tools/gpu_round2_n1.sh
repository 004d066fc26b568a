#!/bin/bash
# Single-GPU round trip on the shipped defaults (three steps per pass, carry in
# shared memory, one-slot TMA ring): the whole -m gpu suite, bench.py at the
# default and at the driver's K, ncu of both headline kernels, launch list,
# a last sweep around the default.  gpurun --timeout 2400 -- bash tools/gpu_round2_n1.sh [tag]
tag=${1:-r2d}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
L=pylabolt_b200/lib
V=$L/variants
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $out/${tag}_pytest_all.log
el "pytest all: $(tail -1 $out/${tag}_pytest_all.log)"
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
el bench
timeout 400 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_k20.json 2> $out/${tag}_bench_k20.err
el bench-k20
PLB_MRT_GENERAL=1 timeout 300 python bench.py --no-extras --no-cpu-baseline > $out/${tag}_bench_mrt_general.json 2> $out/${tag}_bench_mrt_general.err
el bench-general
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
D3=PLB_FUSED_ROWS=64
timeout 700 python tools/fused_sweep.py --models mrt,bgk \
    $L/libplb.so $L/libplb.so:PLB_FUSED_ROWS=128 $L/libplb.so:PLB_FUSED_ROWS=96 $L/libplb.so:PLB_FUSED_ROWS=48 \
    $L/libplb.so:PLB_FUSED_DYNAMIC=1,PLB_FUSED_ROWS=128 \
    $L/libplb.so:PLB_FUSE_DEPTH=2 $L/libplb.so:PLB_FUSE=0 \
    $V/libplb_cb_s2_mb4.so $V/libplb_cb_s2_mb3.so $V/libplb_cb_s1_mb4_d3mb4.so $V/libplb_cb_s1_b64_mb8.so \
    $V/libplb_cb_s1_mb4_late.so $V/libplb_cb_s1_mb4_d3mb2.so $V/libplb_r1.so:PLB_FUSE_DEPTH=2 \
    > $out/${tag}_sweep.txt 2>&1
el sweep
cat $out/${tag}_timeline.log
cut -c1-200 $out/${tag}_sweep.txt
head -c 600 $out/${tag}_bench.json
