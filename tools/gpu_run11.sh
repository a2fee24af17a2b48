set -x
mkdir -p gpurun_out/r2k
python -m pytest tests/test_gpu_level_mode.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/r2k/tests_fast.log 2>&1
tail -3 gpurun_out/r2k/tests_fast.log
for m in 1 0 2; do
SLRBS_PIPE_L2=$m python bench.py --steps 20 --warmup 5 > gpurun_out/r2k/bench_l2mode$m.json 2> gpurun_out/r2k/bench_l2mode$m.err
cut -c1-200 gpurun_out/r2k/bench_l2mode$m.json
done
SLRBS_ROWS_COMPACT=0 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k/bench_rows256.json 2> gpurun_out/r2k/bench_rows256.err
SLRBS_PGS_STAGE=0 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k/bench_nostage.json 2> gpurun_out/r2k/bench_nostage.err
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_pipe -s 203 -c 1 -o gpurun_out/r2k/k_pgs_pipe python bench.py --steps 2 --warmup 3 > gpurun_out/r2k/ncu_f.log 2>&1
SLRBS_PIPE_L2=0 SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_pgs_pipe -s 203 -c 1 -o gpurun_out/r2k/k_pgs_pipe_l2mode0 python bench.py --steps 2 --warmup 3 > gpurun_out/r2k/ncu_f0.log 2>&1
ls -la gpurun_out/r2k
