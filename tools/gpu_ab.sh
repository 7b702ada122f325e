# A/B runs of bench.py on one GPU: each line of ARGS (separated by ';') is one run
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_ab.sh r02d "--volume-layout texture;--volume-layout linear"'
TAG=$1
IFS=';' read -ra RUNS <<< "$2"
mkdir -p gpurun_out
i=0
for r in "${RUNS[@]}"; do
  env $AB_ENV python bench.py --no-cpu --no-gather --steps 16 $r > gpurun_out/${TAG}_ab$i.json 2> gpurun_out/${TAG}_ab$i.err
  tail -c 600 gpurun_out/${TAG}_ab$i.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_ab$i.json").read().strip().splitlines()[-1])
print("RUN $i [$r]", {k: (round(d[k], 4) if isinstance(d[k], float) else d[k]) for k in ("value", "ms_per_step", "n_recomputed_total")})
print("   stages", {k: round(v, 4) for k, v in d["stages_ms_per_step"].items()})
if d.get("e2e"):
    print("   e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v, 4) for k, v in d["e2e"]["stages_ms_per_step"].items()})
PY
  i=$((i+1))
done
