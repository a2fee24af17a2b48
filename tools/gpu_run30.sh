set -x
mkdir -p gpurun_out/r2ae
python -m pytest tests/test_gpu_parity.py tests/test_gpu_broadphase.py tests/test_gpu_full_size.py tests/test_gpu_golden.py tests/test_gpu_extensions.py tests/test_gpu_sdf.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/r2ae/tests_detect.log 2>&1
tail -3 gpurun_out/r2ae/tests_detect.log
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 12 --csv --log-file gpurun_out/r2ae/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2ae/ncu_l.log 2>&1
grep k_detect_rows gpurun_out/r2ae/launches.csv | head -2
