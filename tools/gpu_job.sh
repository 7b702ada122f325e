# what the round-end driver does, in one gpurun call: full GPU parity suite, smoke, default bench line, reference arm,
# and the ncu capture that profiles/roofline_traffic.json is made from (tools/update_roofline_traffic.py, run here later)
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_job.sh r02a'
TAG=${1:-r02}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --maxfail=10 ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "frames_per_sec", "n_recomputed_total", "stages_ms_per_step")})
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "avg_launch_ms")})
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "stages_ms_per_step")} if d.get("e2e") else None)
print("cpu", d.get("cpu_baseline"))
print("gather", d.get("gather"))
PY
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
tail -c 900 gpurun_out/${TAG}_bench_ref.json
BENCH="python bench.py --steps 4 --warmup 3 --timesteps 8 --no-e2e --no-cpu --no-gather"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_trace -f $BENCH > gpurun_out/${TAG}_prof_trace.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
ls -la gpurun_out | tail -12
