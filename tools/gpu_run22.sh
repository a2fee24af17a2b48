set -x
mkdir -p gpurun_out/r2w
python -m pytest tests/test_gpu_level_mode.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_div.py tests/test_gpu_regressions_r2.py -x -q -m gpu > gpurun_out/r2w/tests_fast.log 2>&1
tail -3 gpurun_out/r2w/tests_fast.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2w/bench_default.json 2> gpurun_out/r2w/bench_default.err
cut -c1-200 gpurun_out/r2w/bench_default.json
ls -la gpurun_out/r2w
