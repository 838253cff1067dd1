mkdir -p gpurun_out
( python -m pytest tests/test_gpu_prove.py tests/test_gpu_r1cs.py -x -q ) > gpurun_out/r02x_tests.log 2>&1; tail -3 gpurun_out/r02x_tests.log
python bench/prove_ab.py 22 5 > gpurun_out/r02x_prove.jsonl 2> gpurun_out/r02x.err; cat gpurun_out/r02x_prove.jsonl; tail -2 gpurun_out/r02x.err
