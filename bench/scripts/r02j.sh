mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q ) > gpurun_out/r02j_tests.log 2>&1; tail -4 gpurun_out/r02j_tests.log
python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_TAIL_QUAD=0,1 --steps 5 > gpurun_out/r02j_l21.jsonl 2> gpurun_out/r02j.err; cat gpurun_out/r02j_l21.jsonl
python bench/msm_modes.py --log-n 24 --modes affine --sweep BMPC_TAIL_QUAD=0,1 --steps 3 > gpurun_out/r02j_l24.jsonl 2>> gpurun_out/r02j.err; cat gpurun_out/r02j_l24.jsonl
python bench/msm_modes.py --log-n 24 --modes affine --no-precompute --sweep BMPC_TAIL_QUAD=0,1 --steps 3 > gpurun_out/r02j_l24_plain.jsonl 2>> gpurun_out/r02j.err; cat gpurun_out/r02j_l24_plain.jsonl
python bench/prove_ab.py 22 5 BMPC_TAIL_QUAD=0,1 > gpurun_out/r02j_prove.jsonl 2>> gpurun_out/r02j.err; cat gpurun_out/r02j_prove.jsonl
tail -3 gpurun_out/r02j.err
