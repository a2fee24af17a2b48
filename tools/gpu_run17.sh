set -x
mkdir -p gpurun_out/r2r
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r/smoke.log 2>&1; tail -2 gpurun_out/r2r/smoke.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2r/tests_all.log 2>&1
tail -3 gpurun_out/r2r/tests_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2r/bench_default.json 2> gpurun_out/r2r/bench_default.err
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_pipe -s 203 -c 1 -o gpurun_out/r2r/k_pgs_pipe python bench.py --steps 2 --warmup 3 > gpurun_out/r2r/ncu_f.log 2>&1
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 48 --csv --log-file gpurun_out/r2r/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2r/ncu_l.log 2>&1
python bench.py --workload big_scene --steps 20 --warmup 5 > gpurun_out/r2r/bench_big_scene.json 2> gpurun_out/r2r/bench_big_scene.err
python bench.py --workload mixed_pile --steps 20 --warmup 5 > gpurun_out/r2r/bench_mixed_pile.json 2> gpurun_out/r2r/bench_mixed_pile.err
python bench.py --workload pile_1m --steps 20 --warmup 5 > gpurun_out/r2r/bench_pile_1m.json 2> gpurun_out/r2r/bench_pile_1m.err
ls -la gpurun_out/r2r
