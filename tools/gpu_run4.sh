# memcheck of the new kernels on small cases, default + secondary workloads
set -x
mkdir -p gpurun_out/r2d
python bench.py --steps 20 --warmup 5 > gpurun_out/r2d/bench_default.json 2> gpurun_out/r2d/bench_default.err
for w in big_scene mixed_pile pile marble_box car; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2d/bench_$w.json 2> gpurun_out/r2d/bench_$w.err
done
timeout 900 python bench.py --workload pile_1m --steps 20 --warmup 5 > gpurun_out/r2d/bench_pile_1m.json 2> gpurun_out/r2d/bench_pile_1m.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_regressions_r2.py "tests/test_gpu_level_mode.py::test_level_mode_is_bit_identical_to_sequential" -x -q -k "not swinging" > gpurun_out/r2d/memcheck_level.log 2>&1; echo "rc=$?" >> gpurun_out/r2d/memcheck_level.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_color.py -x -q -k "marble_box-kw0 or joints" > gpurun_out/r2d/memcheck_color.log 2>&1; echo "rc=$?" >> gpurun_out/r2d/memcheck_color.log
ls -la gpurun_out/r2d
