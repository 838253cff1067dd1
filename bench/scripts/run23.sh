# fold kernel + small reduce blocks: parity (multiexp, prove, list_mul_matrix), then the block-size sweep
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_cpp_mirror.py -x -q ) > gpurun_out/t23.log 2>&1; tail -6 gpurun_out/t23.log
for RB in 32 64 128 256; do
for L in 21 24; do
BMPC_REDUCE_BLOCK=$RB python bench.py --log-n $L --no-prove --no-ntt --no-r1cs --no-cpu-baseline > gpurun_out/b23_rb${RB}_l$L.json 2> gpurun_out/b23_rb${RB}_l$L.err
python - <<PY
import json
d=json.loads(open("gpurun_out/b23_rb${RB}_l$L.json").read().strip().splitlines()[-1]); print("RB=$RB L=$L", round(d["value"],1), round(d["ms_per_step"],3), d["kernel_ms"])
PY
done
done
