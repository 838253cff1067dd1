mkdir -p gpurun_out
timeout 300 python bench/msm_modes.py --log-n 21 --modes affine > gpurun_out/r03l_l21.jsonl 2> gpurun_out/r03l.err; cat gpurun_out/r03l_l21.jsonl; tail -3 gpurun_out/r03l.err
timeout 300 python bench/msm_modes.py --log-n 24 --modes affine > gpurun_out/r03l_l24.jsonl 2>> gpurun_out/r03l.err; cat gpurun_out/r03l_l24.jsonl
