# N GPUs (N = $1): the driver's scaling line (MSM 2^24 + sharded prove 2^22 + distributed NTT), then config #5
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 --no-r1cs --no-cpu-baseline > gpurun_out/r02_scaling_n$N.json 2> gpurun_out/r02_scaling_n$N.err
tail -3 gpurun_out/r02_scaling_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_scaling_n$N.json").read().strip().splitlines()[-1]); print("N=$N", round(d["value"],1), round(d["ms_per_step"],3), d["kernel_ms"], "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["serial_value"],1), "checked", d["result_checked"]); print("  prove", {k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","matches_known_dlog_expectation","error","n_gpus")}); print("  ntt_dist", d.get("ntt_dist"))
PY
bash bench/scripts/config5.sh $N 26
