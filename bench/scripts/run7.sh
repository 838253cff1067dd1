python -m pytest tests/test_gpu_prove.py tests/test_gpu_msm.py -x -q > gpurun_out/t_prove.log 2>&1; tail -5 gpurun_out/t_prove.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs > gpurun_out/n1.json 2> gpurun_out/n1.err; tail -2 gpurun_out/n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs > gpurun_out/n2.json 2> gpurun_out/n2.err; tail -5 gpurun_out/n2.err
python - <<'PY'
import json
for f in ("n1","n2"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], d["e2e"]["value"]); print("  prove", {k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","matches_known_dlog_expectation","error","n_gpus")})
    except Exception as e: print(f, "ERR", e)
PY
