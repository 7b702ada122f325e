# quick check on one B200: GPU parity suite + a bench line without the CPU leg
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_quick.sh r02b'
TAG=${1:-r02q}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --maxfail=10 ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log
python bench.py --no-cpu ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1200 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "frames_per_sec", "n_recomputed_total", "stages_ms_per_step")})
print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "frac", "avg_launch_ms", "traffic", "traffic_source")})
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "stages_ms_per_step")} if d.get("e2e") else None)
print("gather", {k: d["gather"][k] for k in ("frames_per_sec", "photon_map_build_ms", "raymarch_ms")} if d.get("gather") else None)
PY
