# launch list of the shard-sized multiexp (2^21 points, reduce blocks of 32)
mkdir -p gpurun_out
BMPC_REDUCE_BLOCK=32 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'msm_|task_|scan_' -c 400 --csv --log-file gpurun_out/l24_launches.csv python bench.py --log-n 21 --steps 3 --warmup 1 --no-prove --no-ntt --no-r1cs --no-cpu-baseline > gpurun_out/l24.log 2>&1
python bench/launch_summary.py gpurun_out/l24_launches.csv gpurun_out/l24_summary.csv | head -30; tail -3 gpurun_out/l24.log | cut -c1-300
