set -x
mkdir -p gpurun_out/r2e
python -m pytest tests/test_gpu_level_mode.py tests/test_gpu_bpp.py tests/test_gpu_sdf.py tests/test_gpu_regressions_r2.py tests/test_gpu_full_size.py -q --maxfail=20 --timeout=900 2>&1 | tail -40 > gpurun_out/r2e/tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2e/bench_gacc.json 2> gpurun_out/r2e/bench_gacc.err
SLRBS_ROWS_COMPACT=1 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e/bench_gacc_compact.json 2> gpurun_out/r2e/bench_gacc_compact.err
SLRBS_PIPE_STREAM=0 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e/bench_gacc_nostream.json 2> gpurun_out/r2e/bench_gacc_nostream.err
SLRBS_PGS_GACC=0 SLRBS_ROWS_COMPACT=1 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e/bench_pipe_compact.json 2> gpurun_out/r2e/bench_pipe_compact.err
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_gacc -s 203 -c 1 -o gpurun_out/r2e/k_pgs_gacc python bench.py --steps 2 --warmup 3 > gpurun_out/r2e/ncu_f.log 2>&1
ls -la gpurun_out/r2e
