# whole-wave job dealing (BMPC_AFF_WHOLE_WAVES) and mixed additions in the bucket reduction: parity + A/B
python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q > gpurun_out/t16.log 2>&1; tail -3 gpurun_out/t16.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs"
$B > gpurun_out/w1.json 2> gpurun_out/w1.err
BMPC_AFF_WHOLE_WAVES=0 $B > gpurun_out/w0.json 2> gpurun_out/w0.err
$B --no-prove --log-n 21 > gpurun_out/w1_21.json 2> gpurun_out/w1_21.err
BMPC_AFF_WHOLE_WAVES=0 $B --no-prove --log-n 21 > gpurun_out/w0_21.json 2> gpurun_out/w0_21.err
$B --no-prove --log-n 22 > gpurun_out/w1_22.json 2> gpurun_out/w1_22.err
BMPC_AFF_WHOLE_WAVES=0 $B --no-prove --log-n 22 > gpurun_out/w0_22.json 2> gpurun_out/w0_22.err
python - <<'PY'
import json
for f in ("w1","w0","w1_21","w0_21","w1_22","w0_22"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], (d.get("prove") or {}).get("all_s"), (d.get("prove") or {}).get("matches_known_dlog_expectation"))
    except Exception as e: print(f, "ERR", e)
PY
tail -2 gpurun_out/w1.err
