#!/bin/bash
# Chunks marching in alternating x directions (PLB_FUSED_ALTERNATE): parity,
# A/B sweep, DRAM traffic of the dominant kernel on both headline workloads.
#   gpurun --timeout 900 -- bash tools/gpu_round2k.sh [tag]
tag=${1:-r2k}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
L=pylabolt_b200/lib
timeout 400 python -m pytest tests/test_gpu_zz_fused_depth4.py tests/test_gpu_zz_fused_depth3.py tests/test_gpu_fused.py tests/test_gpu_full_size.py -m gpu -q 2>&1 | tail -8 > $out/${tag}_pytest.log
el "pytest: $(tail -1 $out/${tag}_pytest.log)"
timeout 300 python tools/fused_sweep.py --models mrt,bgk \
    $L/libplb.so $L/libplb.so:PLB_FUSED_ALTERNATE=0 $L/libplb.so $L/libplb.so:PLB_FUSED_ALTERNATE=0 \
    $L/libplb.so:PLB_FUSED_ROWS=32 $L/libplb.so:PLB_FUSED_ROWS=128 \
    > $out/${tag}_sweep.txt 2>&1
el sweep
for wl in channel cavity; do
  timeout 200 ncu --set full --clock-control none --import-source on \
      -k regex:k_bulk_fused -s 2 -c 1 -f -o $out/${tag}_ncu_${wl}_fused \
      python bench.py --workload $wl --steps 12 --warmup 12 --no-extras --no-cpu-baseline --no-parity > $out/${tag}_ncu_bench_$wl.log 2>&1
  ncu -i $out/${tag}_ncu_${wl}_fused.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_${wl}_fused.csv 2>/dev/null
  ncu -i $out/${tag}_ncu_${wl}_fused.ncu-rep --page details > $out/${tag}_ncu_details_${wl}_fused.txt 2>/dev/null
done
el ncu
cat $out/${tag}_timeline.log
tail -3 $out/${tag}_pytest.log
cut -c1-200 $out/${tag}_sweep.txt
python - $out $tag <<'PY'
import csv, sys
out, tag = sys.argv[1:3]
for wl in ("channel", "cavity"):
    try:
        rows = list(csv.reader(open(f"{out}/{tag}_ncu_raw_{wl}_fused.csv")))
        d = dict(zip(rows[0], rows[-1]))
        print(wl, d["Kernel Name"][:40], "read", d["dram__bytes_read.sum"], "write", d["dram__bytes_write.sum"],
              "ms", d["gpu__time_duration.sum"], "regs", d["launch__registers_per_thread"],
              "fp64", d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"])
    except Exception as e:
        print(wl, "FAILED", e)
PY
