#!/bin/bash
# The shipped tree, one GPU: whole -m gpu suite, smoke(), bench.py (default K and
# the driver's K), launch list.   gpurun --timeout 900 -- bash tools/gpu_round2l.sh [tag]
tag=${1:-r2l}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $out/${tag}_pytest_all.log
el "pytest all: $(tail -1 $out/${tag}_pytest_all.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
el "smoke: $(tail -1 $out/${tag}_smoke.log | cut -c1-160)"
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
el bench
timeout 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_k20.json 2> $out/${tag}_bench_k20.err
el bench-k20
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $out/${tag}_launches_channel.csv \
    python bench.py --steps 12 --warmup 4 --no-extras --no-cpu-baseline --no-parity > /dev/null 2>&1
el launch-list
cat $out/${tag}_timeline.log
for f in bench bench_k20; do
python - $out/${tag}_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "%.2f GLUPS" % d["value"], d["roofline"]["kernel"], "frac %.3f" % d["roofline"]["frac"],
          "traffic", d["roofline"]["traffic"], "e2e %.1f" % d["e2e"]["value"], "parity", d["parity"]["max_rel_err"],
          "clocks", d["clocks"]["sm_mhz"])
    for k, v in (d.get("extra") or {}).items():
        print("   ", k, "%.2f GLUPS" % v["value"], "frac %.3f" % v["roofline"]["frac"], "traffic", v["roofline"]["traffic"],
              "e2e %.1f" % v["e2e"]["value"], "parity", v["parity"]["max_rel_err"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
