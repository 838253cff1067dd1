mkdir -p gpurun_out
( python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_multi.py tests/test_gpu_r1cs.py -x -q ) > gpurun_out/r02w_tests.log 2>&1; tail -2 gpurun_out/r02w_tests.log
python bench/prove_ab.py 22 5 > gpurun_out/r02w_prove.jsonl 2> gpurun_out/r02w.err; cat gpurun_out/r02w_prove.jsonl
python bench/prove_ab.py 22 5 "" 8 > gpurun_out/r02w_prove_share8.jsonl 2>> gpurun_out/r02w.err; cat gpurun_out/r02w_prove_share8.jsonl
tail -2 gpurun_out/r02w.err
