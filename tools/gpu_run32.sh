set -x
mkdir -p gpurun_out/r2ag
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_level_mode.py tests/test_gpu_extensions.py tests/test_gpu_sdf.py -x -q -m gpu > gpurun_out/r2ag/tests.log 2>&1
tail -3 gpurun_out/r2ag/tests.log
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 12 --csv --log-file gpurun_out/r2ag/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2ag/ncu_l.log 2>&1
grep k_contact_rows_lv gpurun_out/r2ag/launches.csv | head -2
