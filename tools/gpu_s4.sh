# session 4, run F: new gather kernel parity + bench
python -m pytest tests/test_gather.py tests/test_bound.py tests/test_tracer.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s4f_pytest.log
tail -5 gpurun_out/s4f_pytest.log
python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu > gpurun_out/s4f_bench.json 2> gpurun_out/s4f_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s4f_bench.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()}, d.get("tests_fetching_voxels"), d["gather"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s4f_bench.err").read()[-1500:])
PY
python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --bound-log2 -1 > gpurun_out/s4f_bench_nb.json 2> gpurun_out/s4f_bench_nb.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s4f_bench_nb.json").read().strip().splitlines()[-1])
    print("no bound:", d["gather"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s4f_bench_nb.err").read()[-1500:])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gather_kernel' -c 1 -o gpurun_out/s4f_gather -f python bench.py --steps 2 --warmup 1 --timesteps 6 --no-e2e --no-cpu > gpurun_out/s4f_ncu.log 2>&1
