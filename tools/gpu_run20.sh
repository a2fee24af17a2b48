set -x
mkdir -p gpurun_out/r2u
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u/smoke.log 2>&1; tail -1 gpurun_out/r2u/smoke.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2u/tests_all.log 2>&1
tail -3 gpurun_out/r2u/tests_all.log
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 48 --csv --log-file gpurun_out/r2u/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2u/ncu_l.log 2>&1
grep -E "k_contact_rows_lv|k_contact_levels" gpurun_out/r2u/launches.csv | head -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r2u/bench_default.json 2> gpurun_out/r2u/bench_default.err
python bench.py --workload mixed_pile --steps 20 --warmup 5 > gpurun_out/r2u/bench_mixed_pile.json 2> gpurun_out/r2u/bench_mixed_pile.err
ls -la gpurun_out/r2u
