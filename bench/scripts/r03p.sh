mkdir -p gpurun_out
BMPC_LIB_PATH=$PWD/bellman_mpc_b200/libexp_g2calls.so timeout 300 python bench/msm_modes.py --group g2 --log-n 21 --modes affine --sweep BMPC_AFF_MINB=3,4,1 > gpurun_out/r03p_calls.jsonl 2> gpurun_out/r03p.err; cat gpurun_out/r03p_calls.jsonl; tail -3 gpurun_out/r03p.err
BMPC_LIB_PATH=$PWD/bellman_mpc_b200/libexp_g2calls.so timeout 300 python bench/msm_modes.py --group g2 --log-n 22 --modes affine > gpurun_out/r03p_calls22.jsonl 2>> gpurun_out/r03p.err; cat gpurun_out/r03p_calls22.jsonl
timeout 300 python bench/msm_modes.py --group g2 --log-n 22 --modes affine > gpurun_out/r03p_base22.jsonl 2>> gpurun_out/r03p.err; cat gpurun_out/r03p_base22.jsonl
