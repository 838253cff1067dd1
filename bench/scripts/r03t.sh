mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "shard" ) > gpurun_out/r03t_tests.log 2>&1; tail -5 gpurun_out/r03t_tests.log
bash bench/scripts/scale3.sh 2; cp gpurun_out/s3_n2.json gpurun_out/r03t_scaling_n2.json
