# session 5, run Z: light-volume ray caster with batched samples
python -m pytest tests/test_raycast.py tests/test_gather.py -m gpu -x -q 2>&1 | tail -2
CPM_RAYCAST_BATCH=2 python -m pytest tests/test_raycast.py -m gpu -x -q 2>&1 | tail -1
for RB in 1 2 4; do
CPM_RAYCAST_BATCH=$RB python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/s5z_bench.json 2> gpurun_out/s5z_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5z_bench.json").read().strip().splitlines()[-1])
    g=d["gather"]; print("RB=$RB raycast_ms", round(g["light_volume_raycast_ms"],3), "view fps", round(d["view_frames_per_sec"],1))
except Exception as e:
    print("failed", e); print(open("gpurun_out/s5z_bench.err").read()[-1500:])
PY
done
