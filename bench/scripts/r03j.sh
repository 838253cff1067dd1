mkdir -p gpurun_out
timeout 300 python bench/msm_modes.py --log-n 24 --modes affine > gpurun_out/r03j_l24.jsonl 2> gpurun_out/r03j.err; cat gpurun_out/r03j_l24.jsonl; tail -3 gpurun_out/r03j.err
timeout 300 python bench/msm_modes.py --log-n 21 --modes affine > gpurun_out/r03j_l21.jsonl 2>> gpurun_out/r03j.err; cat gpurun_out/r03j_l21.jsonl
timeout 300 python bench/msm_modes.py --group g2 --log-n 21 --modes affine > gpurun_out/r03j_g2_l21.jsonl 2>> gpurun_out/r03j.err; cat gpurun_out/r03j_g2_l21.jsonl
python bench/prove_ab.py 22 5 > gpurun_out/r03j_prove.jsonl 2>> gpurun_out/r03j.err; cat gpurun_out/r03j_prove.jsonl | cut -c1-400; tail -2 gpurun_out/r03j.err
( timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r03j_tests.log 2>&1; tail -3 gpurun_out/r03j_tests.log
