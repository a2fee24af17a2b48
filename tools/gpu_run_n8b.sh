set -x
mkdir -p gpurun_out/r2n8b
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612"
for b in 8 6 5; do
  SLRBS_PILE_BLOCK=$b $TR bench.py --gpus 8 --workload pile_1m --steps 40 --warmup 5 > gpurun_out/r2n8b/bench_pile_1m_block$b.json 2> gpurun_out/r2n8b/bench_pile_1m_block$b.err
done
$TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n8b/bench_default.json 2> gpurun_out/r2n8b/bench_default.err
ls -la gpurun_out/r2n8b
