# session 5, run X: selection count read-back overlapped with the opacity-bound refresh
python -m pytest tests/test_sort.py tests/test_host_processors.py tests/test_configs.py tests/test_bound.py tests/test_workspace.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
python bench.py --steps 32 --warmup 3 --no-cpu --no-e2e --no-gather > gpurun_out/s5x_bench.json 2> gpurun_out/s5x_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5x_bench.json").read().strip().splitlines()[-1])
    print(d["value"], round(d["ms_per_step"],4), round(d["wall_ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/s5x_bench.err").read()[-2500:])
PY
done
