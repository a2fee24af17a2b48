set -x
mkdir -p gpurun_out/r2s
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s/smoke.log 2>&1; tail -2 gpurun_out/r2s/smoke.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2s/tests_all.log 2>&1
tail -3 gpurun_out/r2s/tests_all.log
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 12 --csv --log-file gpurun_out/r2s/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2s/ncu_l.log 2>&1
grep k_detect_rows gpurun_out/r2s/launches.csv | head -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2s/bench_default.json 2> gpurun_out/r2s/bench_default.err
ls -la gpurun_out/r2s
