#!/bin/bash
# Final single-GPU evidence of the shipped defaults: the whole -m gpu suite,
# bench.py at the default K and at the driver's K, the nine-rate MRT line, the
# reference arm, ncu of the dominant kernel on both headline workloads, launch list.
#   gpurun --timeout 1800 -- bash tools/gpu_round2g.sh [tag]
tag=${1:-r2g}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $out/${tag}_pytest_all.log
el "pytest all: $(tail -1 $out/${tag}_pytest_all.log)"
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
el bench
timeout 400 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_k20.json 2> $out/${tag}_bench_k20.err
el bench-k20
PLB_MRT_GENERAL=1 timeout 300 python bench.py --no-extras --no-cpu-baseline > $out/${tag}_bench_mrt_general.json 2> $out/${tag}_bench_mrt_general.err
el bench-general
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
el reference-arm
for wl in channel cavity; do
  timeout 240 ncu --set full --clock-control none --import-source on \
      -k regex:k_bulk_fused -s 3 -c 1 -f -o $out/${tag}_ncu_${wl}_fused \
      python bench.py --workload $wl --steps 12 --warmup 12 --no-extras --no-cpu-baseline --no-parity > $out/${tag}_ncu_bench_$wl.log 2>&1
  ncu -i $out/${tag}_ncu_${wl}_fused.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_${wl}_fused.csv 2>/dev/null
  ncu -i $out/${tag}_ncu_${wl}_fused.ncu-rep --page details > $out/${tag}_ncu_details_${wl}_fused.txt 2>/dev/null
done
el ncu
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $out/${tag}_launches_channel.csv \
    python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline --no-parity > /dev/null 2>&1
el launch-list
cat $out/${tag}_timeline.log
tail -2 $out/${tag}_pytest_all.log
for f in bench bench_k20 bench_mrt_general bench_reference_arm; do
python - $out/${tag}_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    p = d.get("parity") or {}
    print(sys.argv[1], "%.2f GLUPS" % d["value"], "frac %.3f" % (d.get("roofline") or {}).get("frac", 0),
          "e2e %.1f" % e.get("value", 0), "parity", p.get("max_rel_err"), "steps", p.get("steps_compared"),
          "clocks", (d.get("clocks") or {}).get("sm_mhz"))
    for k, v in (d.get("extra") or {}).items():
        print("   ", k, "%.2f GLUPS" % v["value"], "frac %.3f" % v["roofline"]["frac"],
              "e2e %.1f" % v["e2e"]["value"], "parity", v["parity"]["max_rel_err"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
