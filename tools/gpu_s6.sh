TAG=${1:-s6f}
bash tools/gpu_job.sh ${TAG}
V=abvariants
bash tools/gpu_ab_env.sh ${TAG} "CPM_B200_LIB=$V/fmad/libcpm_b200.so CPM_HOST_LIB=$V/fmad/libcpm_host.so --;CPM_BOUND_TEXTURE=1 --"
BENCH="python bench.py --steps 4 --warmup 3 --timesteps 8 --no-e2e --no-cpu --no-gather"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:detect_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_detect -f $BENCH > gpurun_out/${TAG}_prof_detect.log 2>&1
