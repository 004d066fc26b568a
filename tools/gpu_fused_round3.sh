#!/bin/bash
# Third (short) GPU round trip: deep-flag prefetch, v1 kernel on the same box.
tag=${1:-h}
out=gpurun_out
mkdir -p $out
L=pylabolt_b200/lib
V=$L/variants
timeout 40 python tools/fused_sweep.py $L/libplb.so > $out/${tag}_sweep.txt 2>&1
timeout 30 python tools/fused_sweep.py --models mrt $V/libplb_v1.so:PLB_FUSED_ROWS=32 >> $out/${tag}_sweep.txt 2>&1
timeout 30 python tools/fused_sweep.py --models mrt $V/libplb_s3_b64_mb6.so >> $out/${tag}_sweep.txt 2>&1
timeout 30 python tools/fused_sweep.py --models bgk $V/libplb_bgk_mb3.so >> $out/${tag}_sweep.txt 2>&1
timeout 30 python tools/fused_sweep.py --models bgk $V/libplb_v1.so:PLB_FUSED_ROWS=32 >> $out/${tag}_sweep.txt 2>&1
timeout 30 python tools/fused_sweep.py --models mrt $V/libplb_s0.so >> $out/${tag}_sweep.txt 2>&1
timeout 30 python tools/fused_sweep.py --models mrt $V/libplb_s2_mb3.so >> $out/${tag}_sweep.txt 2>&1
cat $out/${tag}_sweep.txt
