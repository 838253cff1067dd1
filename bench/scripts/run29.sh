# warp-per-bucket heavy combine: parity (multiexp incl. hot buckets, prove), 2^24 / 2^23 timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q 2>&1 | tail -2
for L in 24 23; do
python bench.py --log-n $L --no-prove --no-ntt --no-r1cs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('L=$L', round(d['value'],1), round(d['ms_per_step'],3), d['kernel_ms'])"
done
