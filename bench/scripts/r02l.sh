mkdir -p gpurun_out
BMPC_TAIL_QUAD=2 python -m pytest tests/test_gpu_msm.py -x -q -k "G2 or g2 or 2-" > gpurun_out/r02l_tests.log 2>&1; tail -3 gpurun_out/r02l_tests.log
python bench/msm_modes.py --log-n 21 --group g2 --modes affine --sweep BMPC_TAIL_QUAD=1,2 --steps 3 > gpurun_out/r02l_g2_l21.jsonl 2> gpurun_out/r02l.err; cat gpurun_out/r02l_g2_l21.jsonl
BMPC_REDUCE_BLOCK=32 python bench/msm_modes.py --log-n 21 --group g2 --modes affine --sweep BMPC_TAIL_QUAD=2 --steps 3 > gpurun_out/r02l_g2_l21_rb32.jsonl 2>> gpurun_out/r02l.err; cat gpurun_out/r02l_g2_l21_rb32.jsonl
python bench/prove_ab.py 22 5 BMPC_TAIL_QUAD=1,2 > gpurun_out/r02l_prove.jsonl 2>> gpurun_out/r02l.err; cat gpurun_out/r02l_prove.jsonl
tail -3 gpurun_out/r02l.err
