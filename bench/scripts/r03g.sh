mkdir -p gpurun_out
timeout 300 python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_REDUCE_BLOCK=0,1 > gpurun_out/r03g_l21.jsonl 2> gpurun_out/r03g.err; cat gpurun_out/r03g_l21.jsonl; tail -3 gpurun_out/r03g.err
timeout 300 python bench/msm_modes.py --log-n 24 --modes affine --sweep BMPC_REDUCE_BLOCK=0,1 > gpurun_out/r03g_l24.jsonl 2>> gpurun_out/r03g.err; cat gpurun_out/r03g_l24.jsonl
timeout 300 python bench/msm_modes.py --group g2 --log-n 21 --modes affine --sweep BMPC_REDUCE_BLOCK=0,1 > gpurun_out/r03g_g2_l21.jsonl 2>> gpurun_out/r03g.err; cat gpurun_out/r03g_g2_l21.jsonl
BMPC_REDUCE_BLOCK=1 timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "affine_radix or xyzz" > gpurun_out/r03g_tests.log 2>&1; tail -3 gpurun_out/r03g_tests.log
