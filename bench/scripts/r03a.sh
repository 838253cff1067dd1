mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "affine_radix" > gpurun_out/r03a_tests.log 2>&1; tail -5 gpurun_out/r03a_tests.log
timeout 300 python bench/msm_modes.py --log-n 24 --modes affine --sweep BMPC_SORT_RADIX=0,1 > gpurun_out/r03a_l24.jsonl 2> gpurun_out/r03a.err; cat gpurun_out/r03a_l24.jsonl; tail -3 gpurun_out/r03a.err
timeout 300 python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_SORT_RADIX=0,1 > gpurun_out/r03a_l21.jsonl 2>> gpurun_out/r03a.err; cat gpurun_out/r03a_l21.jsonl
