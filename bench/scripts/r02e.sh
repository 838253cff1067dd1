# round 2: fold / final kernels with inlined field products (BMPC_TAIL_INLINE) at shard size and at 2^24;
# the new GPU tests; the bench line with the pipelined e2e
mkdir -p gpurun_out
python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_TAIL_INLINE=0,1 --steps 5 > gpurun_out/r02e_tail_l21.jsonl 2> gpurun_out/r02e_tail_l21.err; cat gpurun_out/r02e_tail_l21.jsonl; tail -2 gpurun_out/r02e_tail_l21.err
python bench/msm_modes.py --log-n 21 --group g2 --modes affine --sweep BMPC_TAIL_INLINE=0,1 --steps 3 > gpurun_out/r02e_tail_g2_l21.jsonl 2> gpurun_out/r02e_tail_g2_l21.err; cat gpurun_out/r02e_tail_g2_l21.jsonl; tail -2 gpurun_out/r02e_tail_g2_l21.err
python bench/msm_modes.py --log-n 24 --modes affine --sweep BMPC_TAIL_INLINE=0,1 --steps 3 > gpurun_out/r02e_tail_l24.jsonl 2> gpurun_out/r02e_tail_l24.err; cat gpurun_out/r02e_tail_l24.jsonl; tail -2 gpurun_out/r02e_tail_l24.err
( time python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r02e_new.log 2>&1; tail -4 gpurun_out/r02e_new.log
python bench.py --steps 5 --no-cpu-baseline --no-ntt --no-r1cs --no-prove > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -3 gpurun_out/r02e_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02e_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms"], "e2e", d["e2e"])
PY
