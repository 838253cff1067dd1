mkdir -p gpurun_out
( python -m pytest tests/test_gpu_multi.py tests/test_gpu_cpp_mirror.py -x -q ) > gpurun_out/r02u_tests.log 2>&1; tail -3 gpurun_out/r02u_tests.log
python bench/multi_onecall.py --devices 2 --log-n 20 --log-m 20 --steps 3 > gpurun_out/r02u_onecall.jsonl 2> gpurun_out/r02u.err; cut -c1-200 gpurun_out/r02u_onecall.jsonl; tail -3 gpurun_out/r02u.err
