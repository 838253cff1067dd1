# one full ncu capture of the dominant kernel in its present configuration (512-thread blocks, 2^24)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs --no-prove"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_affine -c 1 -o gpurun_out/prof_affine_d -f $B > gpurun_out/ncu_affine_d.log 2>&1
ncu -i gpurun_out/prof_affine_d.ncu-rep --page details > gpurun_out/r01d_msm_accumulate_affine_g1_ncu_details.txt 2>&1
ncu -i gpurun_out/prof_affine_d.ncu-rep --page raw --csv > gpurun_out/r01d_msm_accumulate_affine_g1_ncu_raw.csv 2>&1
rm -f gpurun_out/prof_affine_d.ncu-rep
grep -E "Duration|DRAM Throughput|Issue Slots Busy|Registers Per Thread|Block Size|Grid Size" gpurun_out/r01d_msm_accumulate_affine_g1_ncu_details.txt | head
