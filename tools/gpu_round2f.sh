#!/bin/bash
# Accuracy of the MRT shortcuts over a few hundred steps (bench.py's in-run
# parity against the oracle), the per-model CTA counts, smoke(), the GPU suite.
#   python tools/build_variants.py mrt_div mrt_pair_old mrt_div_pair_old s2
#   gpurun --timeout 1500 -- bash tools/gpu_round2f.sh [tag]
tag=${1:-r2f}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
L=pylabolt_b200/lib
V=$L/variants
for v in default mrt_div mrt_pair_old mrt_div_pair_old; do
  lib=$V/libplb_$v.so; [ $v = default ] && lib=$L/libplb.so
  PLB_LIB=$PWD/$lib timeout 300 python bench.py --steps 200 --no-extras --no-cpu-baseline \
      > $out/${tag}_parity_$v.json 2> $out/${tag}_parity_$v.err
  python - $out/${tag}_parity_$v.json $v <<'PY' >> $out/${tag}_parity.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-18s %.2f GLUPS  parity %.3e (%d steps)  e2e parity %.3e  e2e %.1f GLUPS" % (
        sys.argv[2], d["value"], d["parity"]["max_rel_err"], d["parity"]["steps_compared"],
        d["parity"]["e2e_max_rel_err"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
el parity
timeout 400 python tools/fused_sweep.py --models mrt,bgk \
    $L/libplb.so $L/libplb.so $V/libplb_s2.so $V/libplb_mrt_div.so $V/libplb_mrt_div_pair_old.so \
    $L/libplb.so:PLB_MRT_GENERAL=1 $L/libplb.so:PLB_FUSED_ROWS=32 $L/libplb.so:PLB_FUSED_ROWS=48 \
    > $out/${tag}_sweep.txt 2>&1
el sweep
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $out/${tag}_pytest_all.log
el "pytest all: $(tail -1 $out/${tag}_pytest_all.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
el "smoke: $(tail -1 $out/${tag}_smoke.log)"
cat $out/${tag}_timeline.log
cat $out/${tag}_parity.txt
cut -c1-220 $out/${tag}_sweep.txt
tail -3 $out/${tag}_pytest_all.log
tail -3 $out/${tag}_smoke.log
