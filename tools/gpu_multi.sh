# multi-GPU check on one box: full GPU suite (incl. tests/test_multigpu.py), then the scaling bench line at N = $1
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_multi.sh 2 r02c'
N=${1:-2}
TAG=${2:-r02m}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --maxfail=10 ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --no-cpu ${BENCH_ARGS} \
    > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
tail -c 2500 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n${N}.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "frames_per_sec", "n_recomputed_total", "stages_ms_per_step", "exchange_check")})
print("parallelism", d["config"]["parallelism"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "h2d_gbs", "stages_ms_per_step", "ingest")} if d.get("e2e") else None)
print("gather", d["gather"]["frames_per_sec"] if d.get("gather") else None)
PY
