# session 5, run O: update_sync splat (kernel patch actually applied), drift test, frame timing
python -m pytest tests/test_detector_splat.py tests/test_host_processors.py tests/test_workspace.py tests/test_gather.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 32 --warmup 3 --no-cpu --no-e2e --no-gather > gpurun_out/s5o_bench_n1.json 2> gpurun_out/s5o_bench_n1.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5o_bench_n1.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["value"], round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/s5o_bench_n1.err").read()[-2500:])
PY
