#!/bin/bash
# Last checks of the shipped tree on one GPU: TMA descriptors on tiny lattices,
# smoke(), compute-sanitizer (memcheck, racecheck) over the several-steps-per-pass
# path, the nine-rate MRT at every depth, bench.py at the driver's K.
#   gpurun --timeout 900 -- bash tools/gpu_round2j.sh [tag]
tag=${1:-r2j}
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*" >> $out/${tag}_timeline.log; }
L=pylabolt_b200/lib
python - > $out/${tag}_tiny.log 2>&1 <<'PY'
import sys
sys.path.insert(0, "tests")
from pylabolt_b200 import capi
for nx, ny in ((1, 1), (2, 3), (3, 70), (70, 3), (17, 129)):
    p = capi.Plb(nx, ny, 1.25)
    p.finalize_geometry()
    p.initialize_pop()
    p.step(5)
    p.sync()
    print(nx, ny, p.fused_info())
    p.close()
print("tiny lattices ok")
PY
el "tiny: $(tail -1 $out/${tag}_tiny.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
el "smoke: $(tail -1 $out/${tag}_smoke.log | cut -c1-150)"
for tool in memcheck racecheck; do
  timeout 280 compute-sanitizer --tool $tool --print-limit 5 \
      python -c "import __graft_entry__ as g; g._smoke_fused()" > $out/${tag}_sanitizer_$tool.log 2>&1
  el "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_sanitizer_$tool.log | tail -1)"
done
timeout 200 python tools/fused_sweep.py --models mrt \
    $L/libplb.so:PLB_MRT_GENERAL=1,PLB_FUSE=0 $L/libplb.so:PLB_MRT_GENERAL=1,PLB_FUSE_DEPTH=2 \
    $L/libplb.so:PLB_MRT_GENERAL=1,PLB_FUSE_DEPTH=3 $L/libplb.so:PLB_MRT_GENERAL=1,PLB_FUSE_DEPTH=4 \
    > $out/${tag}_sweep_nine_rates.txt 2>&1
el sweep
timeout 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_k20.json 2> $out/${tag}_bench_k20.err
el bench
cat $out/${tag}_timeline.log
tail -3 $out/${tag}_tiny.log
cut -c1-200 $out/${tag}_sweep_nine_rates.txt
head -c 300 $out/${tag}_bench_k20.json
