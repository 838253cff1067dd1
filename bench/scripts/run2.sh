python -m pytest tests/test_gpu_msm.py -x -q > gpurun_out/t_msm.log 2>&1; tail -3 gpurun_out/t_msm.log
B="python bench.py --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
export BMPC_ACC_AFFINE=1
$B > gpurun_out/v_b128.json 2> gpurun_out/v_b128.err
BMPC_AFF_BLOCKDIM=32 $B > gpurun_out/v_b32.json 2> gpurun_out/v_b32.err
BMPC_AFF_BLOCKDIM=64 $B > gpurun_out/v_b64.json 2> gpurun_out/v_b64.err
BMPC_AFF_BLOCKDIM=32 BMPC_AFF_KSEL=384 $B > gpurun_out/v_b32k384.json 2> gpurun_out/v_b32k384.err
BMPC_AFF_BLOCKDIM=32 $B --log-n 21 > gpurun_out/v_b32_21.json 2> gpurun_out/v_b32_21.err
BMPC_ACC_AFFINE=0 $B --log-n 21 > gpurun_out/v_x_21.json 2> gpurun_out/v_x_21.err
python - <<'PY'
import json
for f in ("v_b128","v_b32","v_b64","v_b32k384","v_b32_21","v_x_21"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/v_b32.err
