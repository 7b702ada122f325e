#!/bin/bash
# Runs on the GPU box (via gpurun): launch list + full ncu captures of the hot kernels.  Output -> gpurun_out/.
set -u
mkdir -p gpurun_out
K='regex:^(trace|onesweep|histogram|detect|splat|classify|minmax|diff|reduce|seed_streams|fill_u32|scatter_fill|directional|mesh_intersect|uniform2d|cell_range|hash)_kernel'
BENCH="python bench.py --steps 4 --warmup 3 --timesteps 8 --no-e2e --no-cpu"
# every launch of our kernels with its device time (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv \
    --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
# full sets: tracer (skip the first-frame full trace: -s 1), then sort / detector / splat
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 2 \
    -o gpurun_out/prof_trace -f $BENCH > gpurun_out/prof_trace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^(onesweep|histogram|detect|splat)_kernel' -s 8 -c 8 \
    -o gpurun_out/prof_frame -f $BENCH > gpurun_out/prof_frame.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^(onesweep|histogram)_kernel' -s 10 -c 5 \
    -o gpurun_out/prof_sort -f python tools/quickbench.py sort26 > gpurun_out/prof_sort.log 2>&1
ls -la gpurun_out
