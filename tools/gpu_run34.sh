set -x
mkdir -p gpurun_out/r2ai
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_broadphase.py tests/test_gpu_div.py tests/test_gpu_level_mode.py -x -q -m gpu > gpurun_out/r2ai/tests.log 2>&1
tail -3 gpurun_out/r2ai/tests.log
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 12 --csv --log-file gpurun_out/r2ai/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2ai/ncu_l.log 2>&1
grep k_detect_rows gpurun_out/r2ai/launches.csv | head -2
