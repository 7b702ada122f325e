# session 4, run C: trimmed scan loop: parity + SCAN sweep
python -m pytest tests/test_bound.py tests/test_tracer.py tests/test_host_processors.py tests/test_configs.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s4c_pytest.log
tail -3 gpurun_out/s4c_pytest.log
B="python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu --no-gather"
for sc in 4 8 16; do
  CPM_TRACE_SCAN=$sc $B > gpurun_out/s4c_scan$sc.json 2> gpurun_out/s4c_scan$sc.err
done
$B --bound-log2 2 > gpurun_out/s4c_scanb2.json 2> gpurun_out/s4c_scanb2.err
$B --bound-log2 4 > gpurun_out/s4c_scanb4.json 2> gpurun_out/s4c_scanb4.err
for f in 4 8 16 b2 b4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s4c_scan$f.json").read().strip().splitlines()[-1])
    print("scan $f", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()}, d.get("tests_fetching_voxels"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/s4c_scan$f.err").read()[-1500:])
PY
done
