# session 4, run M: N=2 bench under torchrun
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 3 --no-cpu > gpurun_out/s4m_bench_n2.json 2> gpurun_out/s4m_bench_n2.err
tail -c 400 gpurun_out/s4m_bench_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/s4m_bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "fps", d["frames_per_sec"], "n", d["n_gpus"])
print("stages", {k:round(v,4) for k,v in d["stages_ms_per_step"].items()})
print("e2e", d["e2e"])
PY
python -m pytest tests/test_sharding.py -m gpu -x -q 2>&1 | tail -3
