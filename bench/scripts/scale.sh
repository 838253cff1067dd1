N=$1
if [ "$N" = "1" ]; then
python bench.py --gpus 1 --steps 5 --warmup 3 --no-r1cs > gpurun_out/s_n1.json 2> gpurun_out/s_n1.err
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/s_n$N.json 2> gpurun_out/s_n$N.err
fi
tail -3 gpurun_out/s_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/s_n$N.json").read().strip().splitlines()[-1]); print("N=$N", round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", round(d["e2e"]["value"],1)); print("  prove", {k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","matches_known_dlog_expectation","error","n_gpus")})
print("  ", d["roofline_int"]["work"], d["roofline_int"]["frac"])
PY
