# A/B of create_proof at 2^22 under the reduce-block and affine-block knobs; 512-thread affine blocks at 2^21 / 2^24
mkdir -p gpurun_out
for V in "" "BMPC_REDUCE_BLOCK=256" "BMPC_AFF_BLOCKDIM=256" "BMPC_AFF_BLOCKDIM=256 BMPC_REDUCE_BLOCK=256"; do
echo "== prove 2^22 [$V]"; env $V python bench/prove_ab.py 22 8 2>&1 | tail -1
done
for L in 21 24; do
BMPC_AFF_BLOCKDIM=512 python bench.py --log-n $L --no-prove --no-ntt --no-r1cs --no-cpu-baseline > gpurun_out/b27_bd512_l$L.json 2> gpurun_out/b27_bd512_l$L.err
python - <<PY
import json
d=json.loads(open("gpurun_out/b27_bd512_l$L.json").read().strip().splitlines()[-1]); print("BD=512 L=$L", round(d["value"],1), round(d["ms_per_step"],3), d["kernel_ms"])
PY
done
BMPC_AFF_BLOCKDIM=512 python -m pytest tests/test_gpu_msm.py -x -q -k "affine" 2>&1 | tail -2
BMPC_AFF_BLOCKDIM=256 python -m pytest tests/test_gpu_msm.py -x -q -k "affine" 2>&1 | tail -2
