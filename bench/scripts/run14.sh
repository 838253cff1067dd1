B="python bench.py --steps 1 --warmup 3 --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
ncu --set full --clock-control none -k regex:msm_scatter_kernel -c 1 -o gpurun_out/prof_scatter -f $B > gpurun_out/ncu_scatter.log 2>&1
ncu --set full --clock-control none -k regex:msm_count_kernel -c 1 -o gpurun_out/prof_count -f $B > gpurun_out/ncu_count.log 2>&1
tail -1 gpurun_out/ncu_count.log
