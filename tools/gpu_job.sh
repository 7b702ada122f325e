# session 4, run G: gather v2 (96 regs, inner jump loop) parity + bench + ncu
python -m pytest tests/test_gather.py tests/test_bound.py tests/test_tracer.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s4g_pytest.log
tail -3 gpurun_out/s4g_pytest.log
python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu > gpurun_out/s4g_bench.json 2> gpurun_out/s4g_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s4g_bench.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()}, d["gather"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s4g_bench.err").read()[-1500:])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gather_kernel' -c 1 -o gpurun_out/s4g_gather -f python bench.py --steps 2 --warmup 1 --timesteps 6 --no-e2e --no-cpu > gpurun_out/s4g_ncu.log 2>&1
