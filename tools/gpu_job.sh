python -m pytest tests/test_gather.py tests/test_grid.py tests/test_configs.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu > gpurun_out/s4r_bench.json 2> gpurun_out/s4r_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s4r_bench.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],4), {k:v for k,v in d["gather"].items() if k!="note"}, d["view_frames_per_sec"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s4r_bench.err").read()[-1500:])
PY
