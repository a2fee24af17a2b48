set -x
mkdir -p gpurun_out/r2ac
SLRBS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_detect_rows -s 203 -c 1 -o gpurun_out/r2ac/k_detect_rows python bench.py --steps 2 --warmup 3 > gpurun_out/r2ac/ncu_d.log 2>&1
ls -la gpurun_out/r2ac
