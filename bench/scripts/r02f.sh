# round 2: upper bound of what a faster block inversion can give (BMPC_EXP_FAKE_INV: wrong sums, timing
# only), one-wave jobs at shard size, and the launch list of one shard-size multiexp
mkdir -p gpurun_out
python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_EXP_FAKE_INV=0,1 --steps 5 > gpurun_out/r02f_fakeinv_l21.jsonl 2> gpurun_out/r02f.err; cat gpurun_out/r02f_fakeinv_l21.jsonl
python bench/msm_modes.py --log-n 24 --modes affine --sweep BMPC_EXP_FAKE_INV=0,1 --steps 3 > gpurun_out/r02f_fakeinv_l24.jsonl 2>> gpurun_out/r02f.err; cat gpurun_out/r02f_fakeinv_l24.jsonl
python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_AFF_WAVES=1,2,3 --steps 5 > gpurun_out/r02f_waves_l21.jsonl 2>> gpurun_out/r02f.err; cat gpurun_out/r02f_waves_l21.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_l21.csv python bench/msm_modes.py --log-n 21 --modes affine --steps 1 > gpurun_out/r02f_ncu.log 2>&1
python bench/launch_summary.py gpurun_out/r02f_launches_l21.csv 2>/dev/null | tail -25
tail -3 gpurun_out/r02f.err
