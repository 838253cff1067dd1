# state with 256/512-thread batched-affine blocks (G1): full GPU suite + default bench line
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/t28.log 2>&1; grep -E "passed|failed|error" gpurun_out/t28.log | tail -3
python bench.py > gpurun_out/r01d_bench_n1.json 2> gpurun_out/b28.err; tail -2 gpurun_out/b28.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r01d_bench_n1.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print({k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","error","matches_known_dlog_expectation")})
print(d["roofline_int"])
PY
python bench.py --log-n 21 --no-prove --no-ntt --no-r1cs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('L=21', round(d['value'],1), round(d['ms_per_step'],3), d['kernel_ms'])"
