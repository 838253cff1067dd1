# full ncu captures: bucket reduction at shard size (2^19 buckets), the two heaviest kernels of the partition sort at 2^24
mkdir -p gpurun_out
B21="python bench/msm_modes.py --log-n 21 --modes affine --steps 1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_reduce_kernel -s 2 -c 1 -o gpurun_out/r03_reduce_2p19 -f $B21 > gpurun_out/r03h_ncu1.log 2>&1; tail -2 gpurun_out/r03h_ncu1.log
B24="python bench/msm_modes.py --log-n 24 --modes affine --steps 1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_partition -s 2 -c 1 -o gpurun_out/r03_rs_partition -f $B24 > gpurun_out/r03h_ncu2.log 2>&1; tail -2 gpurun_out/r03h_ncu2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_scatter -s 2 -c 1 -o gpurun_out/r03_rs_scatter -f $B24 > gpurun_out/r03h_ncu3.log 2>&1; tail -2 gpurun_out/r03h_ncu3.log
ls -la gpurun_out/*.ncu-rep
