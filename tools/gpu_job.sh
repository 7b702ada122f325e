# session 5, run D: persistent-CTA tracer with LPT groups: parity + A/B
( time python -m pytest tests/test_tracer.py tests/test_bound.py tests/test_configs.py tests/test_host_processors.py tests/test_sharding.py -m gpu -x -q ) > gpurun_out/s5d_pytest.log 2>&1
tail -4 gpurun_out/s5d_pytest.log
for R in 1 0; do
CPM_TRACE_REGROUP=$R python bench.py --steps 16 --warmup 3 --no-cpu --no-e2e --no-gather > gpurun_out/s5d_bench_$R.json 2> gpurun_out/s5d_bench_$R.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5d_bench_$R.json").read().strip().splitlines()[-1])
    print("regroup=$R", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/s5d_bench_$R.err").read()[-1500:])
PY
done
python tools/quickbench.py trace > gpurun_out/s5d_trace_1.log 2>&1; grep "trace I" gpurun_out/s5d_trace_1.log
CPM_TRACE_REGROUP=0 python tools/quickbench.py trace > gpurun_out/s5d_trace_0.log 2>&1; grep "trace I" gpurun_out/s5d_trace_0.log
