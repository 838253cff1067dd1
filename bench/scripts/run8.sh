B="python bench.py --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
for n in 21 22; do
BMPC_ACC_AFFINE=1 $B --group g2 --log-n $n > gpurun_out/g2a_$n.json 2>> gpurun_out/g2.err
BMPC_ACC_AFFINE=0 $B --group g2 --log-n $n > gpurun_out/g2x_$n.json 2>> gpurun_out/g2.err
done
for g in 2 3 4 6; do
BMPC_ACC_AFFINE=1 BMPC_AFF_GMAX=$g $B --log-n 22 > gpurun_out/gm${g}_22.json 2>> gpurun_out/g2.err
done
BMPC_ACC_AFFINE=0 $B --log-n 22 > gpurun_out/gmx_22.json 2>> gpurun_out/g2.err
python - <<'PY'
import json
for f in ("g2a_21","g2x_21","g2a_22","g2x_22","gm2_22","gm3_22","gm4_22","gm6_22","gmx_22"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/g2.err
BMPC_ACC_AFFINE=1 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_affine -c 1 -o gpurun_out/prof_affine2 -f $B --steps 1 --warmup 1 > gpurun_out/ncu_affine2.log 2>&1
tail -2 gpurun_out/ncu_affine2.log
