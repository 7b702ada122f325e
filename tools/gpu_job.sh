# session 5, run K: workspace loader on the GPU + 512^3 grid-kernel ncu capture
python -m pytest tests/test_workspace.py tests/test_host_processors.py tests/test_u3d.py -m gpu -x -q 2>&1 | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:(minmax8|diff8|range8)_kernel' -s 2 -c 16 -o gpurun_out/s5k_grids -f python tools/quickbench.py grids512 > gpurun_out/s5k_ncu_grids.log 2>&1
tail -5 gpurun_out/s5k_ncu_grids.log
