mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "affine_radix or xyzz" > gpurun_out/r03c_tests.log 2>&1; tail -3 gpurun_out/r03c_tests.log
timeout 300 python bench/msm_modes.py --log-n 24 --modes affine --sweep BMPC_SORT_RADIX=0,1 > gpurun_out/r03c_l24.jsonl 2> gpurun_out/r03c.err; cat gpurun_out/r03c_l24.jsonl; tail -3 gpurun_out/r03c.err
BMPC_SORT_RADIX=1 timeout 300 python bench/msm_modes.py --log-n 24 --modes affine --sweep BMPC_RS_CHUNK_LOG=12,13,15,16 > gpurun_out/r03c_l24c.jsonl 2> gpurun_out/r03c.err; cat gpurun_out/r03c_l24c.jsonl; tail -3 gpurun_out/r03c.err
timeout 300 python bench/msm_modes.py --log-n 21 --modes affine --sweep BMPC_SORT_RADIX=0,1 > gpurun_out/r03c_l21.jsonl 2>> gpurun_out/r03c.err; cat gpurun_out/r03c_l21.jsonl
BMPC_SORT_RADIX=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rs_|msm_count|msm_scatter" -c 10 --csv --log-file gpurun_out/r03c_launches.csv python bench/msm_modes.py --log-n 24 --modes affine --steps 1 > gpurun_out/r03c.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r03c_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size') if 'Grid Size' in h else None
for r in rows[1:]:
    print(r[ki][:60], r[gi] if gi is not None else '', r[vi])
PY
