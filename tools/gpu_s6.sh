# session job: GPU parity suite, A/B of variants (abvariants/*) on the resident C4 frame, ncu capture of the tracer
TAG=${1:-s6b}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --maxfail=10 ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -12 gpurun_out/${TAG}_pytest.log
V=abvariants
bash tools/gpu_ab_env.sh ${TAG} "CPM_BOUND_TEXTURE=1 --;CPM_BOUND_TEXTURE=0 --;CPM_B200_LIB=$V/det0/libcpm_b200.so CPM_HOST_LIB=$V/det0/libcpm_host.so --"
BENCH="python bench.py --steps 4 --warmup 3 --timesteps 8 --no-e2e --no-cpu --no-gather"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_trace -f $BENCH > gpurun_out/${TAG}_prof_trace.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:detect_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_detect -f $BENCH > gpurun_out/${TAG}_prof_detect.log 2>&1
ls -la gpurun_out | tail -5
