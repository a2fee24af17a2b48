set -x
mkdir -p gpurun_out/r2l
python -m pytest tests -x -q -m gpu > gpurun_out/r2l/tests_all.log 2>&1
tail -3 gpurun_out/r2l/tests_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2l/bench_default.json 2> gpurun_out/r2l/bench_default.err
SLRBS_PGS_STAGE=0 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l/bench_default_nostage.json 2> gpurun_out/r2l/bench_default_nostage.err
SLRBS_PGS_STAGE=0 python bench.py --workload pile_1m --steps 20 --warmup 5 > gpurun_out/r2l/bench_pile_1m_pipe.json 2> gpurun_out/r2l/bench_pile_1m_pipe.err
SLRBS_PGS_STAGE=1 python bench.py --workload pile_1m --steps 20 --warmup 5 > gpurun_out/r2l/bench_pile_1m_stage.json 2> gpurun_out/r2l/bench_pile_1m_stage.err
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_pipe -s 203 -c 1 -o gpurun_out/r2l/k_pgs_pipe python bench.py --steps 2 --warmup 3 > gpurun_out/r2l/ncu_f.log 2>&1
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_detect_rows -s 203 -c 1 -o gpurun_out/r2l/k_detect_rows python bench.py --steps 2 --warmup 3 > gpurun_out/r2l/ncu_d.log 2>&1
ls -la gpurun_out/r2l
