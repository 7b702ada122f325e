# A/B runs of bench.py on one GPU with different environments: runs separated by ';', each "ENV=.. ENV=.. -- bench args"
TAG=$1
IFS=';' read -ra RUNS <<< "$2"
mkdir -p gpurun_out
i=0
for r in "${RUNS[@]}"; do
  envs="${r%%--*}"; args="${r#*--}"
  env $envs python bench.py --no-cpu --no-gather --no-e2e --steps 16 $args > gpurun_out/${TAG}_ab$i.json 2> gpurun_out/${TAG}_ab$i.err
  tail -c 400 gpurun_out/${TAG}_ab$i.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_ab$i.json").read().strip().splitlines()[-1])
print("RUN $i [$r]", {k: (round(d[k], 4) if isinstance(d[k], float) else d[k]) for k in ("value", "ms_per_step", "n_recomputed_total")}, {k: round(v, 4) for k, v in d["stages_ms_per_step"].items()})
PY
  i=$((i+1))
done
