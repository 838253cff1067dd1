# start-time stagger of the co-resident batched-affine blocks (BMPC_AFF_STAGGER_NS): A/B
python -m pytest tests/test_gpu_msm.py -x -q -k "large or hot" > gpurun_out/t19.log 2>&1; tail -2 gpurun_out/t19.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs --no-prove"
for S in 0 40000 100000 250000; do
  BMPC_AFF_STAGGER_NS=$S $B --log-n 21 > gpurun_out/g21_$S.json 2> gpurun_out/g21_$S.err
  BMPC_AFF_STAGGER_NS=$S $B --log-n 22 > gpurun_out/g22_$S.json 2> gpurun_out/g22_$S.err
done
for S in 0 100000 250000; do
  BMPC_AFF_STAGGER_NS=$S $B > gpurun_out/g24_$S.json 2> gpurun_out/g24_$S.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g2*_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
