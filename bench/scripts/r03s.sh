mkdir -p gpurun_out
timeout 300 python bench/msm_modes.py --group g2 --log-n 21 --modes affine --sweep BMPC_AFF_KSEL=384,128 > gpurun_out/r03s_g2_k.jsonl 2> gpurun_out/r03s.err; cat gpurun_out/r03s_g2_k.jsonl; tail -3 gpurun_out/r03s.err
timeout 300 python bench/msm_modes.py --group g2 --log-n 21 --modes affine --sweep BMPC_AFF_GMAX=4,6,8 >> gpurun_out/r03s_g2_k.jsonl 2>> gpurun_out/r03s.err; tail -3 gpurun_out/r03s_g2_k.jsonl
