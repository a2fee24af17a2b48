set -x
mkdir -p gpurun_out/r2m
for mb in 4 6 8; do
SLRBS_LIB=$PWD/tools/exp/libslrbs_rows$mb.so SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 12 --csv --log-file gpurun_out/r2m/launches_rows$mb.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2m/ncu_rows$mb.log 2>&1
grep k_contact_rows_lv gpurun_out/r2m/launches_rows$mb.csv | head -2
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2m/bench_default.json 2> gpurun_out/r2m/bench_default.err
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 48 --csv --log-file gpurun_out/r2m/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2m/ncu_l.log 2>&1
python bench.py --workload pile_1m --steps 20 --warmup 5 > gpurun_out/r2m/bench_pile_1m.json 2> gpurun_out/r2m/bench_pile_1m.err
python bench.py --workload marble_box --steps 20 --warmup 5 > gpurun_out/r2m/bench_marble_box.json 2> gpurun_out/r2m/bench_marble_box.err
python bench.py --workload car --steps 20 --warmup 5 > gpurun_out/r2m/bench_car.json 2> gpurun_out/r2m/bench_car.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m/bench_reference.json 2> gpurun_out/r2m/bench_reference.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2m/smoke.log 2>&1
ls -la gpurun_out/r2m
