# strong scaling of the 2^24 G1 multiexp and the 2^22 prove on N GPUs of one box (N = $1)
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 --no-ntt --no-r1cs --no-cpu-baseline > gpurun_out/s3_n$N.json 2> gpurun_out/s3_n$N.err
tail -2 gpurun_out/s3_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/s3_n$N.json").read().strip().splitlines()[-1]); print("N=$N", round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", round(d["e2e"]["value"],1)); print("  prove", {k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","matches_known_dlog_expectation","error","n_gpus")})
PY
