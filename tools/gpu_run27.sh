set -x
mkdir -p gpurun_out/r2ab
python -m pytest tests -x -q -m gpu > gpurun_out/r2ab/tests_all.log 2>&1
tail -3 gpurun_out/r2ab/tests_all.log
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 12 --csv --log-file gpurun_out/r2ab/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2ab/ncu_l.log 2>&1
grep k_detect_rows gpurun_out/r2ab/launches.csv | head -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2ab/bench_default.json 2> gpurun_out/r2ab/bench_default.err
cut -c1-200 gpurun_out/r2ab/bench_default.json
