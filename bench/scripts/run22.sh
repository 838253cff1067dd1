# re-entry check of HEAD: full GPU suite, default bench (both arms), shard-sized multiexp probe
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/t22.log 2>&1; tail -4 gpurun_out/t22.log
( time python bench.py ) > gpurun_out/b22.json 2> gpurun_out/b22.err; tail -3 gpurun_out/b22.err; cut -c1-600 gpurun_out/b22.json
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/b22_ref.json 2> gpurun_out/b22_ref.err; tail -3 gpurun_out/b22_ref.err; cut -c1-600 gpurun_out/b22_ref.json
for L in 21 22; do
python bench.py --log-n $L --no-prove --no-ntt --no-r1cs --no-cpu-baseline > gpurun_out/b22_l$L.json 2> gpurun_out/b22_l$L.err
python - <<PY
import json
d=json.loads(open("gpurun_out/b22_l$L.json").read().strip().splitlines()[-1]); print("L=$L", round(d["value"],1), round(d["ms_per_step"],3), d["kernel_ms"], d["roofline_int"]["work"])
PY
done
