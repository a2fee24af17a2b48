set -x
mkdir -p gpurun_out/r2y
python bench.py --steps 20 --warmup 5 > gpurun_out/r2y/bench_20.json 2> gpurun_out/r2y/bench_20.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2y/bench_3.json 2> gpurun_out/r2y/bench_3.err
python bench.py --workload pile_1m --steps 10 --warmup 3 > gpurun_out/r2y/bench_pile.json 2> gpurun_out/r2y/bench_pile.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2y/bench_torchrun1.json 2> gpurun_out/r2y/bench_torchrun1.err
tail -c 300 gpurun_out/r2y/*.err
ls -la gpurun_out/r2y
