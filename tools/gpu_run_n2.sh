# two GPUs: the NCCL paths (tests skipped on a one-GPU box), scene-sharded bench, partitioned pile
set -x
mkdir -p gpurun_out/r2n2
python -m pytest tests/test_gpu_partition.py tests/test_gpu_pile_device.py -q --timeout=900 2>&1 | tail -15 > gpurun_out/r2n2/tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n2/bench_default.json 2> gpurun_out/r2n2/bench_default.err
$TR bench.py --gpus 2 --workload pile --steps 40 --warmup 5 > gpurun_out/r2n2/bench_pile.json 2> gpurun_out/r2n2/bench_pile.err
$TR bench.py --gpus 2 --workload pile_1m --steps 20 --warmup 5 > gpurun_out/r2n2/bench_pile_1m.json 2> gpurun_out/r2n2/bench_pile_1m.err
$TR tests/run_partition_multi.py --steps 40 --device-rebuild 1 --native 0 > gpurun_out/r2n2/pile_torch_exchange.json 2> gpurun_out/r2n2/pile_torch_exchange.err
$TR tests/run_partition_multi.py --steps 40 --device-rebuild 1 --native 1 > gpurun_out/r2n2/pile_native_exchange.json 2> gpurun_out/r2n2/pile_native_exchange.err
ls -la gpurun_out/r2n2
