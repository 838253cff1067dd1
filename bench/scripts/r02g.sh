# round 2: Kaliski inversion on the device (parity first), its effect at shard size and at 2^24, and
# create_proof with one stream per multiexp (BMPC_PROOF_SLOTS=8) on the whole proof and on one rank's share of 8
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_multi.py -x -q ) > gpurun_out/r02g_tests.log 2>&1; tail -4 gpurun_out/r02g_tests.log
python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_AFF_WAVES=1,2 --steps 5 > gpurun_out/r02g_l21.jsonl 2> gpurun_out/r02g.err; cat gpurun_out/r02g_l21.jsonl
python bench/msm_modes.py --log-n 24 --modes affine --steps 3 > gpurun_out/r02g_l24.jsonl 2>> gpurun_out/r02g.err; cat gpurun_out/r02g_l24.jsonl
python bench/msm_modes.py --log-n 21 --group g2 --modes affine --steps 3 > gpurun_out/r02g_g2_l21.jsonl 2>> gpurun_out/r02g.err; cat gpurun_out/r02g_g2_l21.jsonl
python bench/prove_ab.py 22 5 BMPC_PROOF_SLOTS=3,8 > gpurun_out/r02g_prove.jsonl 2>> gpurun_out/r02g.err; cat gpurun_out/r02g_prove.jsonl
python bench/prove_ab.py 22 5 BMPC_PROOF_SLOTS=3,8 8 > gpurun_out/r02g_prove_share8.jsonl 2>> gpurun_out/r02g.err; cat gpurun_out/r02g_prove_share8.jsonl
tail -3 gpurun_out/r02g.err
