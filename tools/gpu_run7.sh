set -x
mkdir -p gpurun_out/r2g
python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 2>&1 | tail -30 > gpurun_out/r2g/tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g/bench_default.json 2> gpurun_out/r2g/bench_default.err
SLRBS_PGS_STAGE=1 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g/bench_stage.json 2> gpurun_out/r2g/bench_stage.err
SLRBS_LEVEL_GROUP=32 python bench.py --workload pile --steps 40 --warmup 5 > gpurun_out/r2g/bench_pile_g32.json 2> gpurun_out/r2g/bench_pile_g32.err
python bench.py --workload pile --steps 40 --warmup 5 > gpurun_out/r2g/bench_pile.json 2> gpurun_out/r2g/bench_pile.err
python bench.py --workload marble_box --steps 20 --warmup 5 > gpurun_out/r2g/bench_marble_box.json 2> gpurun_out/r2g/bench_marble_box.err
ls -la gpurun_out/r2g
