mkdir -p gpurun_out
timeout 300 python bench/msm_modes.py --group g2 --log-n 21 --modes affine,xyzz > gpurun_out/r03o_base.jsonl 2> gpurun_out/r03o.err; cat gpurun_out/r03o_base.jsonl; tail -3 gpurun_out/r03o.err
BMPC_LIB_PATH=$PWD/bellman_mpc_b200/libexp_g2calls.so timeout 300 python bench/msm_modes.py --group g2 --log-n 21 --modes affine,xyzz > gpurun_out/r03o_calls.jsonl 2>> gpurun_out/r03o.err; cat gpurun_out/r03o_calls.jsonl; tail -3 gpurun_out/r03o.err
