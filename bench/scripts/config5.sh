# BASELINE config #5 on N GPUs of one box (N = $1): G1 and G2 multiexp over 2^26 points split across the
# GPUs (known-dlog result check inside bench.py), batch scalar multiplication 2^22 (bench/config5.py)
N=${1:-8}
L=${2:-26}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for G in g1 g2; do
  $TR --master-port 2961$N bench.py --gpus $N --log-n $L --group $G --steps 3 --warmup 3 --no-ntt --no-r1cs --no-cpu-baseline --no-prove --no-plain > gpurun_out/c5_${G}_2p${L}_n$N.json 2> gpurun_out/c5_${G}_2p${L}_n$N.err
  tail -2 gpurun_out/c5_${G}_2p${L}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c5_${G}_2p${L}_n$N.json").read().strip().splitlines()[-1])
    print("$G N=$N", round(d["value"],1), "Mpts/s", round(d["ms_per_step"],2), "ms", d["kernel_ms"], "checked", d["result_checked"], "e2e", round(d["e2e"]["value"],1), "setup_s", d["config"]["setup_s"], "frac", d["roofline"]["frac"] if d.get("roofline") else None)
except Exception as e:
    print("no line:", e)
PY
done
$TR --master-port 2962$N bench/config5.py --gpus $N > gpurun_out/c5_batchmul_n$N.jsonl 2> gpurun_out/c5_batchmul_n$N.err
tail -2 gpurun_out/c5_batchmul_n$N.err; cut -c1-330 gpurun_out/c5_batchmul_n$N.jsonl
