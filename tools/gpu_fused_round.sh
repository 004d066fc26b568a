#!/bin/bash
# One GPU round trip for the two-steps-per-pass path (run under gpurun), most
# important first so that a cut-off call still leaves the essentials in
# gpurun_out/: parity, the A/B bench lines, the variant sweep, one full ncu
# capture of k_bulk_fused2, a full-size parity test, the cavity A/B.
tag=${1:-f}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
L=pylabolt_b200/lib
el start
timeout 240 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
el "pytest: $(tail -1 $out/${tag}_pytest.log)"
timeout 150 python bench.py --steps 100 --no-extras --no-cpu-baseline \
    > $out/${tag}_bench_channel_fused.json 2> $out/${tag}_bench_channel_fused.err
el bench-fused
PLB_FUSE=0 timeout 150 python bench.py --steps 100 --no-extras --no-cpu-baseline \
    > $out/${tag}_bench_channel_unfused.json 2> $out/${tag}_bench_channel_unfused.err
el bench-unfused
timeout 60 python tools/fused_sweep.py $L/libplb.so:PLB_FUSE=0 $L/libplb.so > $out/${tag}_sweep.txt 2>&1
timeout 200 python tools/fused_sweep.py --models mrt \
    $L/libplb.so:PLB_FUSED_ROWS=32 $L/libplb.so:PLB_FUSED_ROWS=128 $L/libplb.so:PLB_FUSED_ROWS=512 \
    $L/variants/libplb_mb3.so $L/variants/libplb_mb5.so $L/variants/libplb_mb4_pf2.so \
    $L/variants/libplb_mb4_pf4.so $L/variants/libplb_mb3_pf2.so $L/variants/libplb_blk64.so \
    $L/variants/libplb_blk256.so >> $out/${tag}_sweep.txt 2>&1
el sweep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_bulk_fused2 -s 3 -c 1 \
    -f -o $out/${tag}_ncu_channel_fused \
    python bench.py --steps 4 --warmup 4 --no-extras --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
ncu -i $out/${tag}_ncu_channel_fused.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_channel_fused.csv 2>/dev/null
ncu -i $out/${tag}_ncu_channel_fused.ncu-rep --page details > $out/${tag}_ncu_details_channel_fused.txt 2>/dev/null
el ncu-full
timeout 240 python -m pytest tests/test_gpu_full_size.py -x -q -k "channel and BGK" 2>&1 | tail -5 > $out/${tag}_pytest_full.log
el "full-size: $(tail -1 $out/${tag}_pytest_full.log)"
timeout 150 python bench.py --workload cavity --steps 50 --no-extras --no-cpu-baseline \
    > $out/${tag}_bench_cavity_fused.json 2> $out/${tag}_bench_cavity_fused.err
el cavity-fused
PLB_FUSE=0 timeout 150 python bench.py --workload cavity --steps 50 --no-extras --no-cpu-baseline \
    > $out/${tag}_bench_cavity_unfused.json 2> $out/${tag}_bench_cavity_unfused.err
el cavity-unfused
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_channel_fused.csv \
    python bench.py --steps 4 --warmup 4 --no-extras --no-cpu-baseline > /dev/null 2>&1
el launch-list
cat $out/${tag}_timeline.log
cat $out/${tag}_sweep.txt
head -c 600 $out/${tag}_bench_channel_fused.json
