# N real GPUs in ONE process: the multi-device tests (peer copies between distinct devices) and the
# one-call multiexp / proof
N=${1:-8}
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_multi.py tests/test_gpu_cpp_mirror.py -x -q ) > gpurun_out/r02t_tests_n$N.log 2>&1; tail -3 gpurun_out/r02t_tests_n$N.log
python bench/multi_onecall.py --devices $N > gpurun_out/r02_onecall_n$N.jsonl 2> gpurun_out/r02_onecall_n$N.err; cat gpurun_out/r02_onecall_n$N.jsonl | cut -c1-400; tail -3 gpurun_out/r02_onecall_n$N.err
