TAG=${1:-s6j}
mkdir -p gpurun_out
BENCH="python bench.py --steps 4 --warmup 3 --timesteps 8 --no-e2e --no-cpu --no-gather"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:splat_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_splat -f $BENCH > gpurun_out/${TAG}_prof_splat.log 2>&1
tail -2 gpurun_out/${TAG}_prof_splat.log
