B="python bench.py --steps 1 --warmup 1 --no-prove --no-cpu-baseline --no-ntt --no-r1cs"
BMPC_ACC_AFFINE=1 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_affine -c 1 -o gpurun_out/prof_affine -f $B > gpurun_out/ncu_affine.log 2>&1
tail -3 gpurun_out/ncu_affine.log
