mkdir -p gpurun_out
BMPC_SORT_RADIX=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rs_|msm_count|msm_scatter" -c 40 --csv --log-file gpurun_out/r03b_launches.csv python bench/msm_modes.py --log-n 24 --modes affine --steps 1 > gpurun_out/r03b.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r03b_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size') if 'Grid Size' in h else None
for r in rows[1:]:
    print(r[ki][:60], r[gi] if gi is not None else '', r[vi])
PY
