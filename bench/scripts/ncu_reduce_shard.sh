mkdir -p gpurun_out
B21="python bench/msm_modes.py --log-n 21 --modes affine --steps 1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_reduce_kernel -s 2 -c 1 -o gpurun_out/r03b_reduce_2p19 -f $B21 > gpurun_out/r03n_ncu1.log 2>&1; tail -2 gpurun_out/r03n_ncu1.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"msm_reduce|msm_fold|msm_final|msm_combine|rs_|scan_|task_|find_heavy" -s 60 -c 40 --csv --log-file gpurun_out/r03n_tail_launches.csv $B21 > gpurun_out/r03n_ncu2.log 2>&1
python bench/launch_summary.py gpurun_out/r03n_tail_launches.csv
