set -x
mkdir -p gpurun_out/r2t
for B in 16 64 148 296; do
for g in 64 32; do
python tests/dev_measure.py --scene mb1k --B $B --nc 4096 --steps 30 --settle 150 --group $g > gpurun_out/r2t/mb1k_B${B}_g$g.log 2>&1
echo "B=$B g=$g: $(grep 'ms/step' gpurun_out/r2t/mb1k_B${B}_g$g.log | tail -1)"
done
done
ls gpurun_out/r2t
python -m pytest tests/test_gpu_level_mode.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/r2t/tests_fast.log 2>&1
tail -3 gpurun_out/r2t/tests_fast.log
SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 12 --csv --log-file gpurun_out/r2t/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2t/ncu_l.log 2>&1
grep k_contact_rows_lv gpurun_out/r2t/launches.csv | head -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2t/bench_default.json 2> gpurun_out/r2t/bench_default.err
