set -x
mkdir -p gpurun_out/r2x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x/smoke.log 2>&1; tail -1 gpurun_out/r2x/smoke.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2x/tests_all.log 2>&1
tail -3 gpurun_out/r2x/tests_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2x/bench_default.json 2> gpurun_out/r2x/bench_default.err
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_pipe -s 203 -c 1 -o gpurun_out/r2x/k_pgs_pipe python bench.py --steps 2 --warmup 3 > gpurun_out/r2x/ncu_f.log 2>&1
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 48 --csv --log-file gpurun_out/r2x/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2x/ncu_l.log 2>&1
python bench.py --workload pile_1m --steps 20 --warmup 5 > gpurun_out/r2x/bench_pile_1m.json 2> gpurun_out/r2x/bench_pile_1m.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x/bench_reference.json 2> gpurun_out/r2x/bench_reference.err
python bench.py > gpurun_out/r2x/bench_noflags.json 2> gpurun_out/r2x/bench_noflags.err
ls -la gpurun_out/r2x
