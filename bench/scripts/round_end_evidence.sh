# round-end evidence on one GPU: default bench line, reference arm, launch list of the same command
mkdir -p gpurun_out
python bench.py > gpurun_out/evidence_bench.json 2> gpurun_out/evidence_bench.err; tail -2 gpurun_out/evidence_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/evidence_ref.json 2> gpurun_out/evidence_ref.err; cut -c1-400 gpurun_out/evidence_ref.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs --no-prove"
$B > gpurun_out/evidence_b2.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/evidence_launches_msm.csv $B > gpurun_out/evidence_ncu.log 2>&1
python bench/launch_summary.py gpurun_out/evidence_launches_msm.csv > gpurun_out/evidence_launch_summary.csv 2>/dev/null; head -14 gpurun_out/evidence_launch_summary.csv
python - <<'PY'
import json
d=json.loads(open("gpurun_out/evidence_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"))
print(d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["share_of_step"])
print({k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","error","matches_known_dlog_expectation")})
print([ (x["log_m"], round(x["ms"],3)) for x in d.get("ntt",{}).get("sweep",[])])
PY
# launch list of two 2^22 proofs
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/evidence_launches_prove.csv python bench/prove_profile.py 22 > gpurun_out/evidence_ncu_prove.log 2>&1
python bench/launch_summary.py gpurun_out/evidence_launches_prove.csv > gpurun_out/evidence_prove_launch_summary.csv 2>/dev/null; head -12 gpurun_out/evidence_prove_launch_summary.csv
