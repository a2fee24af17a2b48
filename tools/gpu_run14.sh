set -x
mkdir -p gpurun_out/r2o
python -m pytest tests/test_gpu_level_mode.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_div.py -x -q -m gpu > gpurun_out/r2o/tests_fast.log 2>&1
tail -3 gpurun_out/r2o/tests_fast.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2o/bench_default.json 2> gpurun_out/r2o/bench_default.err
cut -c1-200 gpurun_out/r2o/bench_default.json
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_pipe -s 203 -c 1 -o gpurun_out/r2o/k_pgs_pipe python bench.py --steps 2 --warmup 3 > gpurun_out/r2o/ncu_f.log 2>&1
ls -la gpurun_out/r2o
