# two-level partition sort (BMPC_SORT_PARTITION): parity incl. heavy bins, A/B at 2^24 / 2^21 / prove
python -m pytest tests/test_gpu_msm.py -x -q > gpurun_out/t18.log 2>&1; tail -5 gpurun_out/t18.log
BMPC_SORT_PARTITION=1 python -m pytest tests/test_gpu_prove.py -x -q > gpurun_out/t18p.log 2>&1; tail -3 gpurun_out/t18p.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs"
$B > gpurun_out/p1.json 2> gpurun_out/p1.err
BMPC_SORT_PARTITION=0 $B > gpurun_out/p0.json 2> gpurun_out/p0.err
$B --no-prove --log-n 21 > gpurun_out/p1_21.json 2> gpurun_out/p1_21.err
BMPC_SORT_PARTITION=0 $B --no-prove --log-n 21 > gpurun_out/p0_21.json 2> gpurun_out/p0_21.err
python - <<'PY'
import json
for f in ("p1","p0","p1_21","p0_21"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], (d.get("prove") or {}).get("all_s"), (d.get("prove") or {}).get("matches_known_dlog_expectation"))
    except Exception as e: print(f, "ERR", e)
PY
tail -2 gpurun_out/p1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l24_r18.csv $B --no-prove --steps 1 --warmup 1 > gpurun_out/ncu_l24.log 2>&1
