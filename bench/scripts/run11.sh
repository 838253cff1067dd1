python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_r1cs.py -x -q > gpurun_out/t11.log 2>&1; tail -4 gpurun_out/t11.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs > gpurun_out/p2.json 2> gpurun_out/p2.err; tail -2 gpurun_out/p2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/p2.json").read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"]); print("  prove", {k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","error","gpu_launches_per_proof")})
PY
