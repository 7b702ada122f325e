# session 5, run Q (8 GPUs): scaling check of the whole bench line
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 16 --warmup 3 > gpurun_out/s5q_bench_n8.json 2> gpurun_out/s5q_bench_n8.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5q_bench_n8.json").read().strip().splitlines()[-1])
    g=d["gather"]
    print(d["n_gpus"], d["value"], round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print("sharded", g["photon_sharded"]["frames_per_sec"], g["photon_sharded"]["frame_ms"], "replicated", g["replicated_map"]["frames_per_sec"], g["replicated_map"]["frame_ms"], g["replicated_map"]["photon_allgather_ms"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s5q_bench_n8.err").read()[-2500:])
PY
