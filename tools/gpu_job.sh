# round-1 final state: full GPU parity suite, smoke, default bench line, reference arm
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1
tail -5 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
    print(d["value"], round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()})
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print("gather", d["gather"]["frames_per_sec"], d["gather"]["raymarch_ms"], d["gather"]["light_volume_raycast_ms"], "view", d["view_frames_per_sec"], "cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/final_bench.err").read()[-2500:])
PY
