# full GPU suite on the whole-waves build + knob A/B: one wave of bigger jobs, 3 blocks per SM for G1
python -m pytest tests -m gpu -x -q > gpurun_out/t20.log 2>&1; tail -3 gpurun_out/t20.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs --no-prove"
BMPC_AFF_WAVES=1 $B --log-n 21 > gpurun_out/k21_w1.json 2> gpurun_out/k21_w1.err
BMPC_AFF_WAVES=3 $B --log-n 21 > gpurun_out/k21_w3.json 2> gpurun_out/k21_w3.err
BMPC_AFF_WAVES=1 $B --log-n 22 > gpurun_out/k22_w1.json 2> gpurun_out/k22_w1.err
BMPC_AFF_WAVES=3 $B --log-n 22 > gpurun_out/k22_w3.json 2> gpurun_out/k22_w3.err
BMPC_AFF_MINB=3 $B > gpurun_out/k24_m3.json 2> gpurun_out/k24_m3.err
BMPC_AFF_WAVES=3 $B > gpurun_out/k24_w3.json 2> gpurun_out/k24_w3.err
BMPC_AFF_WAVES=5 $B > gpurun_out/k24_w5.json 2> gpurun_out/k24_w5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/k2*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], d["roofline_int"]["work"][-70:])
    except Exception as e: print(f, "ERR", e)
PY
