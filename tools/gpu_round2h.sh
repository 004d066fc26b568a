#!/bin/bash
# Four steps per pass on the hardware: parity tests, sweep over chunk heights,
# bench.py with its in-run oracle comparison.
#   gpurun --timeout 1500 -- bash tools/gpu_round2h.sh [tag]
tag=${1:-r2h}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
L=pylabolt_b200/lib
timeout 600 python -m pytest tests/test_gpu_zz_fused_depth4.py tests/test_gpu_zz_fused_depth3.py -m gpu -q 2>&1 | tail -15 > $out/${tag}_pytest_depth4.log
el "pytest depth 4: $(tail -1 $out/${tag}_pytest_depth4.log)"
timeout 500 python tools/fused_sweep.py --models mrt,bgk \
    $L/libplb.so $L/libplb.so:PLB_FUSE_DEPTH=4 $L/libplb.so:PLB_FUSE_DEPTH=4 \
    $L/libplb.so:PLB_FUSE_DEPTH=4,PLB_FUSED_ROWS=32 $L/libplb.so:PLB_FUSE_DEPTH=4,PLB_FUSED_ROWS=48 \
    $L/libplb.so:PLB_FUSE_DEPTH=4,PLB_FUSED_ROWS=96 $L/libplb.so:PLB_FUSE_DEPTH=4,PLB_FUSED_ROWS=128 \
    $L/libplb.so:PLB_FUSE_DEPTH=4,PLB_MRT_GENERAL=1 \
    > $out/${tag}_sweep.txt 2>&1
el sweep
PLB_FUSE_DEPTH=4 timeout 400 python bench.py --no-cpu-baseline > $out/${tag}_bench_depth4.json 2> $out/${tag}_bench_depth4.err
el bench-depth4
timeout 240 ncu --set full --clock-control none --import-source on \
    -k regex:k_bulk_fused -s 3 -c 1 -f -o $out/${tag}_ncu_channel_fused4 \
    env PLB_FUSE_DEPTH=4 python bench.py --workload channel --steps 8 --warmup 8 --no-extras --no-cpu-baseline --no-parity > $out/${tag}_ncu_bench_channel.log 2>&1
ncu -i $out/${tag}_ncu_channel_fused4.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_channel_fused4.csv 2>/dev/null
ncu -i $out/${tag}_ncu_channel_fused4.ncu-rep --page details > $out/${tag}_ncu_details_channel_fused4.txt 2>/dev/null
el ncu
cat $out/${tag}_timeline.log
tail -4 $out/${tag}_pytest_depth4.log
cut -c1-220 $out/${tag}_sweep.txt
python - $out/${tag}_bench_depth4.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%.2f GLUPS" % d["value"], d["roofline"]["kernel"], "frac %.3f" % d["roofline"]["frac"],
          "e2e %.1f" % d["e2e"]["value"], "parity", d["parity"]["max_rel_err"], d["clocks"])
    for k, v in (d.get("extra") or {}).items():
        print("   ", k, "%.2f GLUPS" % v["value"], v["roofline"]["kernel"], "frac %.3f" % v["roofline"]["frac"],
              "parity", v["parity"]["max_rel_err"])
except Exception as ex:
    print("FAILED", ex); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occ|FP64 is the" $out/${tag}_ncu_details_channel_fused4.txt | cut -c1-150
