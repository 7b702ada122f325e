# session 4, run N: raycast parity + bench
python -m pytest tests/test_raycast.py tests/test_gather.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s4n_pytest.log
tail -5 gpurun_out/s4n_pytest.log
python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu > gpurun_out/s4n_bench.json 2> gpurun_out/s4n_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s4n_bench.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],4), d["gather"], d["view_frames_per_sec"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s4n_bench.err").read()[-1500:])
PY
