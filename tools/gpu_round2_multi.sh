#!/bin/bash
# Multi-GPU round trip of the next round: gpurun --gpus N -- bash tools/gpu_round2_multi.sh N
# (N = 2: parity of both face transports incl. the fused path; any N: scaling lines)
n=${1:-2}
out=gpurun_out
mkdir -p $out
if [ "$n" = "2" ] || [ "$n" = "4" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -8 > $out/m${n}_pytest.log
fi
for wl in channel cavity; do
  timeout 300 python bench.py --gpus $n --workload $wl --steps 100 > $out/m${n}_${wl}_fused.json 2> $out/m${n}_${wl}_fused.err
  PLB_FUSE=0 timeout 300 python bench.py --gpus $n --workload $wl --steps 100 > $out/m${n}_${wl}_unfused.json 2> $out/m${n}_${wl}_unfused.err
done
PLB_FACE=nccl timeout 300 python bench.py --gpus $n --steps 100 > $out/m${n}_channel_fused_nccl.json 2>/dev/null
cat $out/m${n}_pytest.log 2>/dev/null
for f in $out/m${n}_*.json; do echo $f; head -c 200 $f; echo; done
