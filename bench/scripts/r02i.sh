# 2 GPUs: the driver's own scaling line at N=2 (MSM 2^24 + sharded prove with the shared H pipeline), then
# BASELINE config #5 at N=2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 --no-ntt --no-r1cs --no-cpu-baseline > gpurun_out/r02i_n2.json 2> gpurun_out/r02i_n2.err
tail -3 gpurun_out/r02i_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02i_n2.json").read().strip().splitlines()[-1]); print("N=2", round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["serial_value"],1), "checked", d["result_checked"]); print("  prove", {k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","matches_known_dlog_expectation","error","n_gpus","h_pipeline")})
PY
BMPC_H_SPLIT=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-ntt --no-r1cs --no-cpu-baseline > gpurun_out/r02i_n2_nosplit.json 2> gpurun_out/r02i_n2_nosplit.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02i_n2_nosplit.json").read().strip().splitlines()[-1]); print("N=2 H replicated: prove", {k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","matches_known_dlog_expectation","error")})
PY
bash bench/scripts/config5.sh 2 26
