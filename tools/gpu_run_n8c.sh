set -x
mkdir -p gpurun_out/r2n8c
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613"
$TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n8c/bench_default.json 2> gpurun_out/r2n8c/bench_default.err
$TR bench.py --gpus 8 --workload pile_1m --steps 40 --warmup 5 > gpurun_out/r2n8c/bench_pile_1m.json 2> gpurun_out/r2n8c/bench_pile_1m.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29614"
$TR4 bench.py --gpus 4 --workload pile_1m --steps 40 --warmup 5 > gpurun_out/r2n8c/bench_pile_1m_gpus4.json 2> gpurun_out/r2n8c/bench_pile_1m_gpus4.err
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615"
$TR2 bench.py --gpus 2 --workload pile_1m --steps 40 --warmup 5 > gpurun_out/r2n8c/bench_pile_1m_gpus2.json 2> gpurun_out/r2n8c/bench_pile_1m_gpus2.err
$TR2 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n8c/bench_default_gpus2.json 2> gpurun_out/r2n8c/bench_default_gpus2.err
python -m pytest tests -x -q -m gpu -k "partition or pile" > gpurun_out/r2n8c/tests_multi.log 2>&1
tail -3 gpurun_out/r2n8c/tests_multi.log
ls -la gpurun_out/r2n8c
