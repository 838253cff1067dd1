python -m pytest tests/test_gpu_msm.py -x -q > gpurun_out/t_msm.log 2>&1; tail -3 gpurun_out/t_msm.log
B="python bench.py --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
export BMPC_ACC_AFFINE=1
$B > gpurun_out/w_k128.json 2> gpurun_out/w_k128.err
BMPC_AFF_KSEL=384 $B > gpurun_out/w_k384.json 2> gpurun_out/w_k384.err
$B --log-n 21 > gpurun_out/w_21.json 2> gpurun_out/w_21.err
BMPC_AFF_KSEL=384 $B --log-n 21 > gpurun_out/w_21k384.json 2> gpurun_out/w_21k384.err
BMPC_ACC_AFFINE=0 $B --log-n 21 > gpurun_out/w_x21.json 2> gpurun_out/w_x21.err
BMPC_ACC_AFFINE=0 $B > gpurun_out/w_x24.json 2> gpurun_out/w_x24.err
$B --group g2 --log-n 22 > gpurun_out/w_g2.json 2> gpurun_out/w_g2.err
BMPC_ACC_AFFINE=0 $B --group g2 --log-n 22 > gpurun_out/w_g2x.json 2> gpurun_out/w_g2x.err
python - <<'PY'
import json
for f in ("w_k128","w_k384","w_21","w_21k384","w_x21","w_x24","w_g2","w_g2x"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/w_k128.err
