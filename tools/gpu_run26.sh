set -x
mkdir -p gpurun_out/r2aa
for v in base prep4 prep5; do
L=$PWD/tools/exp/libslrbs_$v.so; [ $v = base ] && L=$PWD/slrbs_b200/libslrbs_b200.so
SLRBS_LIB=$L SLRBS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 12 --csv --log-file gpurun_out/r2aa/launches_$v.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2aa/ncu_$v.log 2>&1
grep k_prepare gpurun_out/r2aa/launches_$v.csv | head -2
done
