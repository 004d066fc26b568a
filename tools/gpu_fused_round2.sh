#!/bin/bash
# Second GPU round trip for the two-steps-per-pass path: prefetch-ring variants.
tag=${1:-g}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
L=pylabolt_b200/lib
V=$L/variants
el start
timeout 120 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
el "pytest: $(tail -1 $out/${tag}_pytest.log)"
timeout 150 python tools/fused_sweep.py --models mrt \
    $L/libplb.so:PLB_FUSE=0 $L/libplb.so $L/libplb.so:PLB_FUSED_ROWS=16 $L/libplb.so:PLB_FUSED_ROWS=64 \
    $V/libplb_s0.so $V/libplb_s2_mb3.so $V/libplb_s2_b64.so $V/libplb_s3_b64_mb7.so \
    $V/libplb_s3_b64_mb6.so $V/libplb_s4_b64_mb5.so $V/libplb_s4_b32_mb12.so \
    $V/libplb_s3_b64_mb7.so:PLB_FUSED_ROWS=64 > $out/${tag}_sweep.txt 2>&1
timeout 30 python tools/fused_sweep.py --models bgk $L/libplb.so:PLB_FUSE=0 $L/libplb.so >> $out/${tag}_sweep.txt 2>&1
el sweep
timeout 100 python bench.py --steps 100 --no-extras --no-cpu-baseline \
    > $out/${tag}_bench_channel_fused.json 2> $out/${tag}_bench_channel_fused.err
el bench-channel
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_bulk_fused2 -s 3 -c 1 \
    -f -o $out/${tag}_ncu_channel_fused \
    python bench.py --steps 4 --warmup 4 --no-extras --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
ncu -i $out/${tag}_ncu_channel_fused.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_channel_fused.csv 2>/dev/null
ncu -i $out/${tag}_ncu_channel_fused.ncu-rep --page details > $out/${tag}_ncu_details_channel_fused.txt 2>/dev/null
el ncu-full
timeout 100 python bench.py --workload cavity --steps 50 --no-extras --no-cpu-baseline \
    > $out/${tag}_bench_cavity_fused.json 2> $out/${tag}_bench_cavity_fused.err
el bench-cavity
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_channel_fused.csv \
    python bench.py --steps 4 --warmup 4 --no-extras --no-cpu-baseline > /dev/null 2>&1
el launch-list
cat $out/${tag}_timeline.log
cat $out/${tag}_sweep.txt
