set -x
mkdir -p gpurun_out/r2aj
python -m pytest tests -x -q -m gpu > gpurun_out/r2aj/tests_all.log 2>&1
tail -3 gpurun_out/r2aj/tests_all.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2aj/bench_50.json 2> gpurun_out/r2aj/bench_50.err
