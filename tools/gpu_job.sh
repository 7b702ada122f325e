# session 5, run S: compute-sanitizer memcheck over the kernels written this session
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_grid.py tests/test_bound.py tests/test_gather.py tests/test_raycast.py tests/test_detector_splat.py -m gpu -x -q -k "minmax or diff or value_range or strips or linear or splat_update or raycast_matches or gather_raymarch" > gpurun_out/s5s_memcheck.log 2>&1
echo "memcheck exit: $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/s5s_memcheck.log | head -20
