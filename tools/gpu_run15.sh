set -x
mkdir -p gpurun_out/r2p
python -m pytest tests -x -q -m gpu > gpurun_out/r2p/tests_all.log 2>&1
tail -3 gpurun_out/r2p/tests_all.log
for p in 1 2 3 4 6; do python tests/dev_overlap.py --parts $p --steps 30 > gpurun_out/r2p/overlap_parts$p.log 2>&1; tail -1 gpurun_out/r2p/overlap_parts$p.log; done
for c in 2 3 6 8; do SLRBS_E2E_CHUNKS=$c python bench.py --steps 10 --warmup 3 > gpurun_out/r2p/bench_e2e_chunks$c.json 2> gpurun_out/r2p/bench_e2e_chunks$c.err; done
ls -la gpurun_out/r2p
