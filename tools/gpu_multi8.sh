# 8-GPU check: the multi-GPU worker test (world 4) and the scaling bench line at N = 8
TAG=${1:-r02g}
mkdir -p gpurun_out
( time python -m pytest tests/test_multigpu.py -m gpu -q ) > gpurun_out/${TAG}_pytest_mgpu.log 2>&1
tail -8 gpurun_out/${TAG}_pytest_mgpu.log
for N in 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --no-cpu ${BENCH_ARGS} \
    > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
tail -c 1500 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n${N}.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "frames_per_sec", "n_recomputed_total", "stages_ms_per_step", "exchange_check")})
print("parallelism", d["config"]["parallelism"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "h2d_gbs", "stages_ms_per_step")} if d.get("e2e") else None)
print("gather", d["gather"]["frames_per_sec"] if d.get("gather") else None)
PY
done
