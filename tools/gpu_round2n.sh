#!/bin/bash
# CTA layout of the fused kernel: y-adjacent strips of one chunk (PLB_FUSED_ALTERNATE=1,
# shipped) against x-adjacent chunks of one strip (=3): GLUPS, DRAM bytes, parity.
tag=${1:-r2n}
out=gpurun_out
mkdir -p $out
L=pylabolt_b200/lib
timeout 200 python tools/fused_sweep.py --models mrt,bgk \
    $L/libplb.so $L/libplb.so:PLB_FUSED_ALTERNATE=3 $L/libplb.so $L/libplb.so:PLB_FUSED_ALTERNATE=3 \
    $L/libplb.so:PLB_FUSED_ALTERNATE=2 > $out/${tag}_sweep.txt 2>&1
for alt in 1 3; do
  for wl in channel cavity; do
    PLB_FUSED_ALTERNATE=$alt timeout 100 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none -k regex:k_bulk_fused -s 2 -c 1 --csv --log-file $out/${tag}_dram_${wl}_alt$alt.csv \
        python bench.py --workload $wl --steps 12 --warmup 12 --no-extras --no-cpu-baseline --no-parity > /dev/null 2>&1
  done
done
PLB_FUSED_ALTERNATE=3 timeout 200 python -m pytest tests/test_gpu_zz_fused_depth4.py tests/test_gpu_zz_fused_depth3.py tests/test_gpu_full_size.py -m gpu -q 2>&1 | tail -3 > $out/${tag}_pytest_alt3.log
cut -c1-200 $out/${tag}_sweep.txt
grep -h "dram__\|gpu__time" $out/${tag}_dram_*.csv | cut -d, -f5,13- | cut -c1-160
tail -2 $out/${tag}_pytest_alt3.log
