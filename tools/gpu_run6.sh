set -x
mkdir -p gpurun_out/r2f
python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 2>&1 | tail -40 > gpurun_out/r2f/tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f/bench_stage.json 2> gpurun_out/r2f/bench_stage.err
SLRBS_PGS_STAGE=0 SLRBS_ROWS_COMPACT=1 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f/bench_pipe_compact.json 2> gpurun_out/r2f/bench_pipe_compact.err
SLRBS_E2E_CHUNKS=4 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f/bench_stage_chunks4.json 2> gpurun_out/r2f/bench_stage_chunks4.err
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 48 --csv --log-file gpurun_out/r2f/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2f/ncu_l.log 2>&1
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_stage -s 203 -c 1 -o gpurun_out/r2f/k_pgs_stage python bench.py --steps 2 --warmup 3 > gpurun_out/r2f/ncu_f.log 2>&1
ls -la gpurun_out/r2f
