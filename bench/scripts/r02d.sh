# round 2: new entry points (async waiter, multi-device context, verifier acceptance), then the whole suite
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_prove.py tests/test_gpu_cpp_mirror.py -x -q ) > gpurun_out/r02d_new.log 2>&1; tail -15 gpurun_out/r02d_new.log
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02d_tests.log 2>&1; tail -5 gpurun_out/r02d_tests.log
python bench.py --steps 3 --no-cpu-baseline --no-ntt --no-r1cs --no-prove > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -3 gpurun_out/r02d_bench.err; head -c 400 gpurun_out/r02d_bench.json
