#!/bin/bash
# Final short GPU round trip of round 1: the shipped default (ring of 2, 3 CTAs/SM).
tag=${1:-k}
out=gpurun_out
mkdir -p $out
timeout 40 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -3 > $out/${tag}_pytest.log
timeout 40 python bench.py --steps 100 --no-extras --no-cpu-baseline \
    > $out/${tag}_bench_channel_fused.json 2> $out/${tag}_bench_channel_fused.err
timeout 20 python tools/fused_sweep.py --models bgk pylabolt_b200/lib/libplb.so > $out/${tag}_sweep.txt 2>&1
cat $out/${tag}_pytest.log $out/${tag}_sweep.txt
head -c 300 $out/${tag}_bench_channel_fused.json
