# full GPU test suite, A/B benches of the solver variants, launch list, one full ncu capture of the default solver kernel
set -x
mkdir -p gpurun_out/r2c
python -m pytest tests -m gpu -q --maxfail=30 -x --timeout=900 2>&1 | tail -60 > gpurun_out/r2c/tests.log
python -m pytest tests -m gpu -q --maxfail=40 --timeout=900 --deselect tests/test_gpu_level_mode.py 2>&1 | tail -80 > gpurun_out/r2c/tests_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c/bench_pipe.json 2> gpurun_out/r2c/bench_pipe.err
SLRBS_PGS_PIPE=0 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c/bench_lv.json 2> gpurun_out/r2c/bench_lv.err
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 48 --csv --log-file gpurun_out/r2c/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2c/ncu_l.log 2>&1
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_pipe -s 203 -c 1 -o gpurun_out/r2c/k_pgs_pipe python bench.py --steps 2 --warmup 3 > gpurun_out/r2c/ncu_f.log 2>&1
ls -la gpurun_out/r2c
