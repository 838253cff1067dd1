mkdir -p gpurun_out
python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_EXP_REDUCE_SADD=0,1,2 --steps 5 > gpurun_out/r02k_l21.jsonl 2> gpurun_out/r02k.err; cat gpurun_out/r02k_l21.jsonl
BMPC_REDUCE_BLOCK=64 python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_EXP_REDUCE_SADD=0,1 --steps 5 > gpurun_out/r02k_l21_rb64.jsonl 2>> gpurun_out/r02k.err; cat gpurun_out/r02k_l21_rb64.jsonl
python bench/msm_modes.py --log-n 24 --modes affine --sweep BMPC_EXP_REDUCE_SADD=0,1 --steps 3 > gpurun_out/r02k_l24.jsonl 2>> gpurun_out/r02k.err; cat gpurun_out/r02k_l24.jsonl
python bench/msm_modes.py --log-n 21 --group g2 --modes affine --sweep BMPC_EXP_REDUCE_SADD=0,1 --steps 3 > gpurun_out/r02k_g2_l21.jsonl 2>> gpurun_out/r02k.err; cat gpurun_out/r02k_g2_l21.jsonl
tail -2 gpurun_out/r02k.err
