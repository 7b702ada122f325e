BENCH="python bench.py --steps 4 --warmup 3 --timesteps 8 --no-e2e --no-cpu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -o gpurun_out/prof_trace2 -f $BENCH > gpurun_out/prof_trace2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^(onesweep)_kernel' -s 10 -c 2 -o gpurun_out/prof_sort2 -f python tools/quickbench.py sort26 > gpurun_out/prof_sort2.log 2>&1
