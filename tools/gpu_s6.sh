# session job: parity of the wavefront tracer + A/B on the resident C4 frame
TAG=${1:-s6d}
mkdir -p gpurun_out
( python -m pytest tests/test_bound.py tests/test_tracer.py -m gpu -q --maxfail=5 ) > gpurun_out/${TAG}_pytest_a.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_a.log
( CPM_TRACE_WAVEFRONT=8 python -m pytest tests/test_bound.py tests/test_tracer.py tests/test_host_processors.py tests/test_configs.py -m gpu -q --maxfail=5 ) > gpurun_out/${TAG}_pytest_b.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_b.log
bash tools/gpu_ab_env.sh ${TAG} "CPM_TRACE_WAVEFRONT=0 --;CPM_TRACE_WAVEFRONT=8 --;CPM_TRACE_WAVEFRONT=4 --;CPM_TRACE_WAVEFRONT=12 --;CPM_TRACE_WAVEFRONT=16 --;CPM_TRACE_WAVEFRONT=8 CPM_BOUND_TEXTURE=0 --;CPM_TRACE_WAVEFRONT=8 CPM_TRACE_SCAN=4 --;CPM_TRACE_WAVEFRONT=8 CPM_TRACE_SCAN=16 --"
BENCH="python bench.py --steps 4 --warmup 3 --timesteps 8 --no-e2e --no-cpu --no-gather"
CPM_TRACE_WAVEFRONT=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_walk -f $BENCH > gpurun_out/${TAG}_prof_walk.log 2>&1
CPM_TRACE_WAVEFRONT=8 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
