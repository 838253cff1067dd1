mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_prove.py tests/test_gpu_ntt.py tests/test_gpu_multi.py -x -q -k "not sweep" ) > gpurun_out/r02h_tests.log 2>&1; tail -6 gpurun_out/r02h_tests.log
python bench/prove_ab.py 22 5 BMPC_PROOF_SLOTS=0 > gpurun_out/r02h_prove.jsonl 2> gpurun_out/r02h.err; cat gpurun_out/r02h_prove.jsonl; tail -2 gpurun_out/r02h.err
