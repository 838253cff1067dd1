# full ncu capture (with source) of the dominant kernel in its round-2 state, 2^24
mkdir -p gpurun_out
B="python bench/msm_modes.py --log-n 24 --modes affine --steps 1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_affine -s 2 -c 1 -o gpurun_out/r02_acc_affine_g1 -f $B > gpurun_out/r02p_ncu.log 2>&1
tail -3 gpurun_out/r02p_ncu.log
ls -la gpurun_out/r02_acc_affine_g1.ncu-rep
B2="python bench/msm_modes.py --log-n 21 --group g2 --modes affine --steps 1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_affine -s 2 -c 1 -o gpurun_out/r02_acc_affine_g2 -f $B2 > gpurun_out/r02p_ncu_g2.log 2>&1
tail -2 gpurun_out/r02p_ncu_g2.log
