# Evidence for this state (reduce blocks of 32 + fold kernel + list_mul_matrix): full GPU suite, the
# default bench line of both arms (no profiler), then launch lists and one full ncu capture of the
# dominant kernel of the same command.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/t25.log 2>&1; tail -5 gpurun_out/t25.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01d_bench_reference_arm.json 2> gpurun_out/b25_ref.err
python bench.py > gpurun_out/r01d_bench_n1.json 2> gpurun_out/b25.err; tail -2 gpurun_out/b25.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r01d_bench_n1.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print({k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","error","matches_known_dlog_expectation")})
print(d.get("cpu_baseline"))
PY
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs --no-prove"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_msm_launches.csv $B > gpurun_out/ncu_lm.log 2>&1
python bench/launch_summary.py gpurun_out/r01d_msm_launches.csv gpurun_out/r01d_msm_launch_summary.csv | head -12
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01d_prove_launches.csv python bench/prove_profile.py 22 > gpurun_out/ncu_lp.log 2>&1
python bench/launch_summary.py gpurun_out/r01d_prove_launches.csv gpurun_out/r01d_prove_launch_summary.csv | head -24
ncu --set full --clock-control none --import-source on -k regex:msm_reduce_kernel -c 1 -o gpurun_out/prof_reduce -f $B > gpurun_out/ncu_reduce.log 2>&1
ncu -i gpurun_out/prof_reduce.ncu-rep --page details > gpurun_out/r01d_msm_reduce_g1_ncu_details.txt 2>&1
ncu -i gpurun_out/prof_reduce.ncu-rep --page raw --csv > gpurun_out/r01d_msm_reduce_g1_ncu_raw.csv 2>&1
rm -f gpurun_out/prof_reduce.ncu-rep
