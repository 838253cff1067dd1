mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "affine_radix" > gpurun_out/r03e_tests.log 2>&1; tail -3 gpurun_out/r03e_tests.log
BMPC_SORT_RADIX=1 timeout 300 python bench/msm_modes.py --log-n 24 --modes affine --sweep BMPC_RS_CHUNK_LOG=12,13,14 > gpurun_out/r03e_l24c.jsonl 2> gpurun_out/r03e.err; cat gpurun_out/r03e_l24c.jsonl; tail -3 gpurun_out/r03e.err
BMPC_SORT_RADIX=1 timeout 300 python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_RS_CHUNK_LOG=12,14 > gpurun_out/r03e_l21.jsonl 2>> gpurun_out/r03e.err; cat gpurun_out/r03e_l21.jsonl
BMPC_RS_CHUNK_LOG=12 BMPC_SORT_RADIX=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"rs_|msm_count|msm_scatter" -c 5 --csv --log-file gpurun_out/r03e_launches.csv python bench/msm_modes.py --log-n 24 --modes affine --steps 1 > gpurun_out/r03e.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r03e_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); mi=h.index('Metric Name')
for r in rows[1:]:
    print(r[ki][:40], r[mi], r[vi])
PY
