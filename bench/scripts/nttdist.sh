# distributed four-step EvaluationDomain::fft over N GPUs (N = $1): the bench line's ntt_dist object
N=${1:-2}
mkdir -p gpurun_out
for LOG in 26 28; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 2 --warmup 3 --no-prove --no-r1cs --no-cpu-baseline --ntt-max-log 16 --ntt-dist-log $LOG > gpurun_out/nd_n${N}_l$LOG.json 2> gpurun_out/nd_n${N}_l$LOG.err
tail -1 gpurun_out/nd_n${N}_l$LOG.err | cut -c1-200
python - <<PY
import json
d=json.loads(open("gpurun_out/nd_n${N}_l$LOG.json").read().strip().splitlines()[-1]); print("N=$N", d.get("ntt_dist"))
PY
done
