out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py -q 2>&1 | tail -12 > $out/s2k_pytest.log
for spec in "8 channel p2p" "4 channel p2p" "8 cavity p2p" "8 cavity nccl"; do
  set -- $spec
  PLB_FACE=$3 timeout 300 python bench.py --gpus $1 --workload $2 --steps 100 > $out/s2k_${2}_${3}_n$1.json 2>> $out/s2k.err
done
cat $out/s2k_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s2k_*_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,round(d["value"],2),round(d["ms_per_step"],4),round(d["roofline"]["frac"],4),d["roofline"]["face_transport"],round(d["roofline"]["kernel_share_of_step"],4),round(d["e2e"]["value"],1),d["clocks"])
    except Exception as e:
        print(f,"FAILED",e)
PY
tail -5 $out/s2k.err
