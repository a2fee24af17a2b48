set -x
mkdir -p gpurun_out/r2af
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2af/smoke.log 2>&1; tail -1 gpurun_out/r2af/smoke.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2af/tests_all.log 2>&1
tail -3 gpurun_out/r2af/tests_all.log
python bench.py > gpurun_out/r2af/bench_noflags.json 2> gpurun_out/r2af/bench_noflags.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2af/bench_20.json 2> gpurun_out/r2af/bench_20.err
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 48 --csv --log-file gpurun_out/r2af/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2af/ncu_l.log 2>&1
