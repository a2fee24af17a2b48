# eight GPUs: scene-sharded bench (e2e with every rank copying at once), the 10^6-sphere pile
set -x
mkdir -p gpurun_out/r2n8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612"
$TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n8/bench_default.json 2> gpurun_out/r2n8/bench_default.err
$TR bench.py --gpus 8 --workload pile_1m --steps 30 --warmup 5 > gpurun_out/r2n8/bench_pile_1m.json 2> gpurun_out/r2n8/bench_pile_1m.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613"
$TR4 bench.py --gpus 4 --workload pile_1m --steps 30 --warmup 5 > gpurun_out/r2n8/bench_pile_1m_n4.json 2> gpurun_out/r2n8/bench_pile_1m_n4.err
nvidia-smi topo -m > gpurun_out/r2n8/topo.txt 2>&1
ls -la gpurun_out/r2n8
