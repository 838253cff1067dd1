mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_prove_launches.csv python bench/prove_profile.py 22 > gpurun_out/r02v_ncu.log 2>&1
python bench/launch_summary.py gpurun_out/r02_prove_launches.csv gpurun_out/r02_prove_launch_summary.csv | head -40
