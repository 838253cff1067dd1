B="python bench.py --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/l21_c.csv $B --log-n 21 --steps 2 --warmup 1 > gpurun_out/ncu_l21.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/l24_c.csv $B --log-n 24 --steps 2 --warmup 1 > gpurun_out/ncu_l24.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msm_reduce_kernel -c 1 -o gpurun_out/prof_reduce2 -f $B --log-n 21 --steps 1 --warmup 1 > gpurun_out/ncu_reduce2.log 2>&1
tail -2 gpurun_out/ncu_reduce2.log
