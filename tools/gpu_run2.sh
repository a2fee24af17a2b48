set -x
mkdir -p gpurun_out/r2b
python -m pytest tests/test_gpu_level_mode.py tests/test_gpu_regressions_r2.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -15 > gpurun_out/r2b/tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b/bench_pipe.json 2> gpurun_out/r2b/bench_pipe.err
SLRBS_PGS_PIPE=0 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b/bench_lv.json 2> gpurun_out/r2b/bench_lv.err
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_pipe -s 203 -c 1 -o gpurun_out/r2b/k_pgs_pipe python bench.py --steps 2 --warmup 3 > gpurun_out/r2b/ncu_f.log 2>&1
ls -la gpurun_out/r2b
