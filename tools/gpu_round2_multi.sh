#!/bin/bash
# Multi-GPU round trip: gpurun --gpus N -- bash tools/gpu_round2_multi.sh N [steps]
# N = 2 / 4: parity of both face transports incl. the fused path (pytest, log kept);
# any N: one bench.py line (both headline configs, each with its in-bench
# oracle parity) on the shipped path, plus the unfused and NCCL-face variants.
n=${1:-2}
steps=${2:-100}
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/m${n}_topo.txt 2>&1
if [ "$n" = "2" ] || [ "$n" = "4" ]; then
  timeout 900 python -m pytest tests/test_gpu_multirank.py -q -rA 2>&1 | tail -70 > $out/m${n}_pytest.log
fi
timeout 420 python bench.py --gpus $n --steps $steps > $out/m${n}_fused.json 2> $out/m${n}_fused.err
if [ -z "$MULTI_ONLY_DEFAULT" ]; then
  PLB_FUSE=0 timeout 420 python bench.py --gpus $n --steps $steps > $out/m${n}_unfused.json 2> $out/m${n}_unfused.err
  PLB_FACE=nccl timeout 420 python bench.py --gpus $n --steps $steps > $out/m${n}_fused_nccl.json 2> $out/m${n}_fused_nccl.err
fi
tail -5 $out/m${n}_pytest.log 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/m*_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    o = next(iter(d.get("extra", {}).values()), {})
    print(f, "N=%d" % d["n_gpus"], "%.1f GLUPS" % d["value"], "parity", d["parity"] and d["parity"]["max_rel_err"],
          "| other: %.1f GLUPS" % o.get("value", 0), "parity", o.get("parity") and o["parity"]["max_rel_err"],
          "| e2e %.1f" % d["e2e"]["value"])
PY
