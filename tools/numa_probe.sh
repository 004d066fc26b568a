#!/bin/bash
# 8-GPU box: what the platform says about NUMA, then the strong-scaling cavity
# with and without binding each rank to its GPU's socket.
out=gpurun_out
lscpu | grep -i "numa\|socket\|model name" 
nvidia-smi topo -m 2>/dev/null | head -14
for i in 0 1 2 3 4 5 6 7; do
  id=$(nvidia-smi -i $i --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//')
  echo "gpu $i $id numa $(cat /sys/bus/pci/devices/$id/numa_node 2>/dev/null)"
done
for bind in 1 0; do
  PLB_DEBUG=1 PLB_NUMA_BIND=$bind timeout 300 python bench.py --gpus 8 --workload cavity --steps 50 > $out/s2o_cavity_bind${bind}_n8.json 2> $out/s2o_bind${bind}.err
  grep "\[plb\]" $out/s2o_cavity_bind${bind}_n8.json $out/s2o_bind${bind}.err | head -8
  python - <<PY
import json
d=[json.loads(l) for l in open("$out/s2o_cavity_bind${bind}_n8.json") if l.startswith("{")][-1]
print("bind=$bind", round(d["value"],1), round(d["e2e"]["value"],1), d["e2e"]["breakdown_ms"])
PY
done
