python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q > gpurun_out/t_msm.log 2>&1; tail -3 gpurun_out/t_msm.log
B="python bench.py --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
export BMPC_ACC_AFFINE=1 BMPC_AFF_MINB=4
for n in 21 22 23 24; do
$B --log-n $n > gpurun_out/z_k128_$n.json 2>> gpurun_out/z.err
BMPC_AFF_KSEL=384 $B --log-n $n > gpurun_out/z_k384_$n.json 2>> gpurun_out/z.err
done
python - <<'PY'
import json
for n in (21,22,23,24):
  for f in ("z_k128_%d"%n,"z_k384_%d"%n):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/z.err
