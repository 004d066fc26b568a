#!/bin/bash
# One GPU round trip (run under gpurun): GPU tests, the bench line, the ncu
# launch list and one full ncu capture of the dominant kernel per workload.
# Outputs land in gpurun_out/<tag>_*.
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
if [ "$2" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_channel_mrt.csv \
    python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
for wl in channel cavity; do
  # the channel step launches the bulk kernel three times (two slab-edge
  # columns, then the dominant interior launch): capture an interior one
  skip=3; [ $wl = channel ] && skip=8
  ncu --set full --clock-control none --import-source on -k regex:k_bulk -s $skip -c 1 \
      -f -o $out/${tag}_ncu_${wl} \
      python bench.py --workload $wl --steps 3 --warmup 3 --no-extras --no-cpu-baseline >> $out/${tag}_ncu_bench.log 2>&1
  ncu -i $out/${tag}_ncu_${wl}.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_${wl}.csv 2>/dev/null
  ncu -i $out/${tag}_ncu_${wl}.ncu-rep --page details > $out/${tag}_ncu_details_${wl}.txt 2>/dev/null
done
fi
cat $out/${tag}_pytest.log
head -c 1500 $out/${tag}_bench.json
