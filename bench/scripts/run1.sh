python -m pytest tests/test_gpu_msm.py -x -q > gpurun_out/t_msm.log 2>&1; tail -3 gpurun_out/t_msm.log
B="python bench.py --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
BMPC_ACC_AFFINE=0 $B > gpurun_out/ba0.json 2> gpurun_out/ba0.err
BMPC_ACC_AFFINE=1 $B > gpurun_out/ba1.json 2> gpurun_out/ba1.err
BMPC_ACC_AFFINE=1 BMPC_AFF_GMAX=4 $B > gpurun_out/ba1g4.json 2> gpurun_out/ba1g4.err
BMPC_ACC_AFFINE=1 $B --log-n 21 > gpurun_out/ba1_21.json 2> gpurun_out/ba1_21.err
BMPC_ACC_AFFINE=0 $B --log-n 21 > gpurun_out/ba0_21.json 2> gpurun_out/ba0_21.err
python - <<'PY'
import json
for f in ("ba0","ba1","ba1g4","ba0_21","ba1_21"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/ba1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l21_new.csv $B --log-n 21 --steps 2 --warmup 1 > gpurun_out/ncu_l21.log 2>&1
