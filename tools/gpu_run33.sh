set -x
mkdir -p gpurun_out/r2ah
python -m pytest tests -x -q -m gpu > gpurun_out/r2ah/tests_all.log 2>&1
tail -3 gpurun_out/r2ah/tests_all.log
python bench.py > gpurun_out/r2ah/bench_noflags.json 2> gpurun_out/r2ah/bench_noflags.err
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 48 --csv --log-file gpurun_out/r2ah/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2ah/ncu_l.log 2>&1
