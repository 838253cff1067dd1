python -m pytest tests/ -m gpu -x -q > gpurun_out/t15.log 2>&1; tail -3 gpurun_out/t15.log
B="python bench.py --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
for n in 21 24; do $B --log-n $n > gpurun_out/r_$n.json 2>> gpurun_out/r.err; done
BMPC_ACC_AFFINE=0 $B --log-n 24 > gpurun_out/r_x24.json 2>> gpurun_out/r.err
$B --group g2 --log-n 22 > gpurun_out/r_g2.json 2>> gpurun_out/r.err
python - <<'PY'
import json
for f in ("r_21","r_24","r_x24","r_g2"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
python bench/prove_ab.py 22 6 2>&1 | tail -1
