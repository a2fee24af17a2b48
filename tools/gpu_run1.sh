set -x
mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a/tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench_split.json 2> gpurun_out/r2a/bench_split.err
SLRBS_PGS_SPLIT=0 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench_nosplit.json 2> gpurun_out/r2a/bench_nosplit.err
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2a/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2a/ncu_l.log 2>&1
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_sp -s 210 -c 1 -o gpurun_out/r2a/k_pgs_sp python bench.py --steps 2 --warmup 3 > gpurun_out/r2a/ncu_f.log 2>&1
ls -la gpurun_out/r2a
