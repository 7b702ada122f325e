# round-1 checkpoint: full GPU test pass, smoke, default bench, launch list + full ncu captures -> gpurun_out/r01d_*
set -u
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r01d_pytest.log; tail -3 gpurun_out/r01d_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r01d_bench_c4_n1.json 2> gpurun_out/r01d_bench.err; tail -c 300 gpurun_out/r01d_bench.err
BENCH="python bench.py --steps 4 --warmup 3 --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01d_launches_c4.csv $BENCH > gpurun_out/r01d_launches_bench.log 2>&1
B2="python bench.py --steps 4 --warmup 3 --timesteps 8 --no-e2e --no-cpu"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^(trace|detect|splat|select)_kernel' -s 6 -c 4 -o gpurun_out/r01d_frame -f $B2 > gpurun_out/r01d_frame.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^(gather|raycast)_kernel' -c 2 -o gpurun_out/r01d_view -f $B2 > gpurun_out/r01d_view.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^(range|bound|minmax|diff)_kernel' -c 4 -o gpurun_out/r01d_grids -f $B2 > gpurun_out/r01d_grids.log 2>&1
ls -la gpurun_out | grep r01d
