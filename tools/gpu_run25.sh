set -x
mkdir -p gpurun_out/r2z
python bench.py > gpurun_out/r2z/bench_noflags.json 2> gpurun_out/r2z/bench_noflags.err
cut -c1-300 gpurun_out/r2z/bench_noflags.json
