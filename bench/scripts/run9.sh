python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q > gpurun_out/t9.log 2>&1; tail -3 gpurun_out/t9.log
B="python bench.py --steps 3 --warmup 3 --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
for n in 21 22 23 24; do
$B --log-n $n > gpurun_out/q_$n.json 2>> gpurun_out/q.err
done
python - <<'PY'
import json
for n in (21,22,23,24):
  for f in ("q_%d"%n,):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], d["roofline_int"]["work"][:60])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/q.err
