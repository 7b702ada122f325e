# session 5, run W: full GPU parity suite, smoke, default bench line (round-1 final state)
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s5w_pytest.log 2>&1
tail -5 gpurun_out/s5w_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s5w_smoke.log 2>&1; tail -2 gpurun_out/s5w_smoke.log
python bench.py > gpurun_out/s5w_bench.json 2> gpurun_out/s5w_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5w_bench.json").read().strip().splitlines()[-1])
    print(d["value"], round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()})
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], {k:round(v,3) for k,v in d["e2e"]["stages_ms_per_step"].items()})
    print({k:(round(v["achieved_gbs"]), round(v["frac_of_hbm_peak"],3)) for k,v in d["e2e"]["grid_kernels"].items()})
    print("gather", d["gather"]["frames_per_sec"], d["gather"]["raymarch_ms"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s5w_bench.err").read()[-2500:])
PY
