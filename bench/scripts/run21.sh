# distributed four-step transform: parity on one GPU (emulated worlds), then the full GPU suite
python -m pytest tests/test_gpu_dist_ntt.py -x -q > gpurun_out/t21a.log 2>&1; tail -15 gpurun_out/t21a.log
python -m pytest tests -m gpu -x -q > gpurun_out/t21.log 2>&1; tail -3 gpurun_out/t21.log
