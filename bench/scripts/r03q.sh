mkdir -p gpurun_out
timeout 300 python bench/msm_modes.py --group g2 --log-n 21 --modes affine --sweep BMPC_AFF_BLOCKDIM=128,256 > gpurun_out/r03q_g2_21.jsonl 2> gpurun_out/r03q.err; cat gpurun_out/r03q_g2_21.jsonl; tail -3 gpurun_out/r03q.err
timeout 300 python bench/msm_modes.py --group g2 --log-n 22 --modes affine --sweep BMPC_AFF_BLOCKDIM=128,256 > gpurun_out/r03q_g2_22.jsonl 2>> gpurun_out/r03q.err; cat gpurun_out/r03q_g2_22.jsonl
timeout 300 python bench/msm_modes.py --group g2 --log-n 19 --modes affine,xyzz > gpurun_out/r03q_g2_19.jsonl 2>> gpurun_out/r03q.err; cat gpurun_out/r03q_g2_19.jsonl
python bench/prove_ab.py 22 5 > gpurun_out/r03q_prove.jsonl 2>> gpurun_out/r03q.err; cat gpurun_out/r03q_prove.jsonl | cut -c1-300; tail -2 gpurun_out/r03q.err
