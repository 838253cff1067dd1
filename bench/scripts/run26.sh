# G2 back to 256-thread reduce blocks without fold; 256-thread blocks for the batched-affine kernel (experiment)
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q ) > gpurun_out/t26.log 2>&1; tail -4 gpurun_out/t26.log
python bench.py > gpurun_out/r01d_bench_n1.json 2> gpurun_out/b26.err; tail -2 gpurun_out/b26.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r01d_bench_n1.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print({k:v for k,v in d.get("prove",{}).items() if k in ("value","all_s","error","matches_known_dlog_expectation")})
PY
for BD in 128 256; do
for L in 21 24; do
BMPC_AFF_BLOCKDIM=$BD python bench.py --log-n $L --no-prove --no-ntt --no-r1cs --no-cpu-baseline > gpurun_out/b26_bd${BD}_l$L.json 2> gpurun_out/b26_bd${BD}_l$L.err
python - <<PY
import json
d=json.loads(open("gpurun_out/b26_bd${BD}_l$L.json").read().strip().splitlines()[-1]); print("BD=$BD L=$L", round(d["value"],1), round(d["ms_per_step"],3), d["kernel_ms"])
PY
done
done
