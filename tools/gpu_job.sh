# session 4, run K: clearance with warp vote
python -m pytest tests/test_bound.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s4k_pytest.log
tail -3 gpurun_out/s4k_pytest.log
python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu --no-gather > gpurun_out/s4k_bench.json 2> gpurun_out/s4k_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s4k_bench.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/s4k_bench.err").read()[-1500:])
PY
