set -x
mkdir -p gpurun_out/r2h
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 48 --csv --log-file gpurun_out/r2h/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2h/ncu_l.log 2>&1
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_pipe -s 203 -c 1 -o gpurun_out/r2h/k_pgs_pipe python bench.py --steps 2 --warmup 3 > gpurun_out/r2h/ncu_f.log 2>&1
SLRBS_GRAPH=0 ncu --set full --clock-control none -k regex:k_contact_rows_lv -s 203 -c 1 -o gpurun_out/r2h/k_contact_rows_lv python bench.py --steps 2 --warmup 3 > gpurun_out/r2h/ncu_rows.log 2>&1
python tests/dev_measure.py --scene marble_box --B 16 --nc 1024 --solver 3 --steps 5 --settle 30 > gpurun_out/r2h/bpp_marble_box.log 2>&1
python tests/dev_measure.py --scene car --B 65536 --nc 8 --solver 3 --steps 10 --settle 50 > gpurun_out/r2h/bpp_car.log 2>&1
SLRBS_PILE_BLOCK=6 python bench.py --workload pile_1m --steps 20 --warmup 5 > gpurun_out/r2h/bench_pile_1m_block6.json 2> gpurun_out/r2h/bench_pile_1m_block6.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h/bench_default.json 2> gpurun_out/r2h/bench_default.err
ls -la gpurun_out/r2h
