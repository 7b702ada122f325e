# session 5, run Y: photon-map cell size sweep for the gather
for S in 1.0 1.34 1.5 2.0; do
python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --gather-grid-scale $S > gpurun_out/s5y_bench.json 2> gpurun_out/s5y_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5y_bench.json").read().strip().splitlines()[-1])
    g=d["gather"]; print("scale=$S", g["grid"], round(g["photon_map_build_ms"],3), round(g["raymarch_ms"],3), round(g["frames_per_sec"],1), g["pixels_lit"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s5y_bench.err").read()[-1500:])
PY
done
