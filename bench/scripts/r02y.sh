mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; tail -2 gpurun_out/r02y_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02y_ref.json 2> gpurun_out/r02y_ref.err; cut -c1-300 gpurun_out/r02y_ref.json; tail -2 gpurun_out/r02y_ref.err
