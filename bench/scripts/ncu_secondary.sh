# ncu --set full of the secondary kernels: batch scalar multiplication (mpc.rs:647-706) and the R1CS SpMV
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none -k regex:batch_scalar_mul_kernel -s 1 -c 1 -o gpurun_out/r03_batch_scalar_mul -f python bench/config5.py --log-n 20 --steps 1 > gpurun_out/ncu_sec1.log 2>&1; tail -1 gpurun_out/ncu_sec1.log
timeout 200 ncu --set full --clock-control none -k regex:spmv -s 1 -c 1 -o gpurun_out/r03_r1cs_spmv -f python bench/r1cs_bench.py 18 > gpurun_out/ncu_sec2.log 2>&1; tail -1 gpurun_out/ncu_sec2.log
ls -la gpurun_out/r03_batch_scalar_mul.ncu-rep gpurun_out/r03_r1cs_spmv.ncu-rep
