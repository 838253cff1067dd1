# round 2: suite at HEAD, then the ncu evidence VERDICT r01 item 5/12 asks for: ntt_pass at 2^22 / 2^24 (full
# set: achieved DRAM GB/s, fmaheavy %), the G2 accumulate kernel
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02c_tests.log 2>&1; tail -3 gpurun_out/r02c_tests.log
for L in 22 24; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 3 -c 3 -o gpurun_out/r02c_ntt_$L -f python bench/ntt_profile.py $L 2 > gpurun_out/r02c_ntt_$L.log 2>&1
  ncu -i gpurun_out/r02c_ntt_$L.ncu-rep --page raw --csv > gpurun_out/r02c_ntt_pass_2p${L}_ncu_raw.csv 2>&1
  ncu -i gpurun_out/r02c_ntt_$L.ncu-rep --page details > gpurun_out/r02c_ntt_pass_2p${L}_ncu_details.txt 2>&1
  rm -f gpurun_out/r02c_ntt_$L.ncu-rep
done
B="python bench.py --group g2 --log-n 22 --steps 1 --warmup 3 --no-cpu-baseline --no-ntt --no-r1cs --no-prove"
$B > gpurun_out/r02c_g2_2p22.json 2> gpurun_out/r02c_g2_2p22.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_affine -s 3 -c 1 -o gpurun_out/r02c_g2acc -f $B > gpurun_out/r02c_g2acc.log 2>&1
ncu -i gpurun_out/r02c_g2acc.ncu-rep --page raw --csv > gpurun_out/r02c_msm_accumulate_affine_g2_ncu_raw.csv 2>&1
ncu -i gpurun_out/r02c_g2acc.ncu-rep --page details > gpurun_out/r02c_msm_accumulate_affine_g2_ncu_details.txt 2>&1
rm -f gpurun_out/r02c_g2acc.ncu-rep
grep -E "Duration|DRAM Throughput|Registers Per" gpurun_out/r02c_*_details.txt | head -30
cat gpurun_out/r02c_g2_2p22.json | head -c 600
