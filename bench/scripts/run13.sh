B="python bench.py --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-ntt --no-r1cs --group g2 --log-n 22"
BMPC_AFF_MINB=3 $B > gpurun_out/g2m3.json 2> gpurun_out/g2m.err
$B > gpurun_out/g2m1.json 2>> gpurun_out/g2m.err
BMPC_ACC_AFFINE=0 $B > gpurun_out/g2mx.json 2>> gpurun_out/g2m.err
python - <<'PY'
import json
for f in ("g2m3","g2m1","g2mx"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -2 gpurun_out/g2m.err
