for S in 4 8 12; do
CPM_GATHER_S=$S python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu > gpurun_out/s4s_bench_$S.json 2> gpurun_out/s4s_bench_$S.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s4s_bench_$S.json").read().strip().splitlines()[-1])
    print("S=$S", round(d["gather"]["raymarch_ms"],3))
except Exception as e:
    print("failed", e); print(open("gpurun_out/s4s_bench_$S.err").read()[-800:])
PY
done
