# Round-end evidence: the default bench line (no profiler), then launch lists and one full ncu
# capture of the dominant kernel of the same commands.
python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; tail -2 gpurun_out/final_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs --no-prove"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_msm_r01c.csv $B > gpurun_out/ncu_lm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_prove_r01c.csv python bench/prove_profile.py 22 > gpurun_out/ncu_lp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_affine -c 1 -o gpurun_out/prof_affine3 -f $B > gpurun_out/ncu_affine3.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_n1.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", d["e2e"]["value"])
print({k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","error","matches_known_dlog_expectation","cpu_baseline")})
print(d.get("cpu_baseline"))
print([ (x["log_m"], round(x["ms"],3)) for x in d.get("ntt",{}).get("sweep",[])])
print(d.get("r1cs"))
PY
