# cp.async operand staging in the batched-affine kernel (BMPC_AFF_STAGED) and grouped scatter atomics
# (BMPC_SCATTER_GROUP): parity, A/B, launch list of a 2^21-point multiexp (the 8-GPU shard size)
python -m pytest tests/test_gpu_msm.py -x -q > gpurun_out/t17.log 2>&1; tail -3 gpurun_out/t17.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs --no-prove"
$B > gpurun_out/s1.json 2> gpurun_out/s1.err
BMPC_AFF_STAGED=0 $B > gpurun_out/s0.json 2> gpurun_out/s0.err
BMPC_SCATTER_GROUP=1 $B > gpurun_out/s1g1.json 2> gpurun_out/s1g1.err
BMPC_SCATTER_GROUP=6 $B > gpurun_out/s1g6.json 2> gpurun_out/s1g6.err
$B --log-n 21 > gpurun_out/s1_21.json 2> gpurun_out/s1_21.err
BMPC_AFF_STAGED=0 $B --log-n 21 > gpurun_out/s0_21.json 2> gpurun_out/s0_21.err
python - <<'PY'
import json
for f in ("s1","s0","s1g1","s1g6","s1_21","s0_21"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -2 gpurun_out/s1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/l21_r17.csv $B --log-n 21 --steps 2 --warmup 1 > gpurun_out/ncu_l21.log 2>&1
python bench/ncu_summary.py gpurun_out/l21_r17.csv 2>/dev/null | head -30
