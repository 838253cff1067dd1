mkdir -p gpurun_out
timeout 300 python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_REDUCE_SLOG_ADD=0,1,2 > gpurun_out/r03k_l21.jsonl 2> gpurun_out/r03k.err; cat gpurun_out/r03k_l21.jsonl; tail -3 gpurun_out/r03k.err
BMPC_REDUCE_SLOG_ADD=1 timeout 300 python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_REDUCE_BLOCK=32,64,128 >> gpurun_out/r03k_l21.jsonl 2> gpurun_out/r03k.err; tail -3 gpurun_out/r03k_l21.jsonl; tail -3 gpurun_out/r03k.err
timeout 300 python bench/msm_modes.py --log-n 24 --modes affine --sweep BMPC_REDUCE_SLOG_ADD=0,1 > gpurun_out/r03k_l24.jsonl 2>> gpurun_out/r03k.err; cat gpurun_out/r03k_l24.jsonl
