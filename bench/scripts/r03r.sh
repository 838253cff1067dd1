mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r03r_tests.log 2>&1; tail -3 gpurun_out/r03r_tests.log
python bench.py > gpurun_out/r03r_bench.json 2> gpurun_out/r03r_bench.err; tail -2 gpurun_out/r03r_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r03r_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"))
print({k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","error","matches_known_dlog_expectation")})
PY
