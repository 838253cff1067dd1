B="python bench.py --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
export BMPC_ACC_AFFINE=1
$B > gpurun_out/y_m1k128.json 2> gpurun_out/y.err
BMPC_AFF_MINB=4 $B > gpurun_out/y_m4k128.json 2>> gpurun_out/y.err
BMPC_AFF_MINB=4 BMPC_AFF_KSEL=384 $B > gpurun_out/y_m4k384.json 2>> gpurun_out/y.err
BMPC_AFF_KSEL=384 $B > gpurun_out/y_m1k384.json 2>> gpurun_out/y.err
for n in 21 22 23; do
BMPC_AFF_MINB=4 $B --log-n $n > gpurun_out/y_m4_$n.json 2>> gpurun_out/y.err
BMPC_ACC_AFFINE=0 $B --log-n $n > gpurun_out/y_x_$n.json 2>> gpurun_out/y.err
done
python - <<'PY'
import json
for f in ("y_m1k128","y_m4k128","y_m4k384","y_m1k384","y_m4_21","y_x_21","y_m4_22","y_x_22","y_m4_23","y_x_23"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/y.err
BMPC_ACC_AFFINE=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/l21_d.csv $B --log-n 21 --steps 2 --warmup 1 > gpurun_out/ncu_l21.log 2>&1
