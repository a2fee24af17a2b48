set -x
mkdir -p gpurun_out/r2ak
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ak/smoke.log 2>&1; tail -1 gpurun_out/r2ak/smoke.log
python -m pytest tests/test_gpu_level_mode.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2ak/tests.log 2>&1; tail -2 gpurun_out/r2ak/tests.log
