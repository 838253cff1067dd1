# final evidence of this state: default bench line of both arms (no profiler), then the launch list of the same command
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01d_bench_reference_arm.json 2> gpurun_out/b30_ref.err
python bench.py > gpurun_out/r01d_bench_n1.json 2> gpurun_out/b30.err; tail -1 gpurun_out/b30.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r01d_bench_n1.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print({k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","error","matches_known_dlog_expectation")})
PY
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_msm_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs --no-prove > gpurun_out/ncu_lm.log 2>&1
python bench/launch_summary.py gpurun_out/r01d_msm_launches.csv gpurun_out/r01d_msm_launch_summary.csv | head -12
