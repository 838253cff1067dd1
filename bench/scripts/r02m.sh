mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02m_tests.log 2>&1; tail -4 gpurun_out/r02m_tests.log
python bench.py > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; tail -2 gpurun_out/r02m_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02m_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["serial_value"],1), "frac", round(d["roofline"]["frac"],3))
print({k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","error","matches_known_dlog_expectation")}, d["prove"].get("cpu_baseline",{}).get("speedup"))
print(d.get("cpu_baseline")); print(d.get("no_precompute"))
print([ (x["log_m"], round(x["ms"],3), x.get("e2e_host_buffer_ms")) for x in d.get("ntt",{}).get("sweep",[])])
print(d.get("r1cs"))
PY
