# session 5, run V (8 GPUs): whole bench line with the peer exchange kernel
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 32 --warmup 3 --check-exchange > gpurun_out/s5v_bench_n8.json 2> gpurun_out/s5v_bench_n8.err
echo "exit $?"
grep -h "exchange check\|unavailable" gpurun_out/s5v_bench_n8.err | head -3
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5v_bench_n8.json").read().strip().splitlines()[-1])
    g=d["gather"]
    print(d["n_gpus"], d["value"], round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print(d["config"]["parallelism"])
    print("sharded", g["photon_sharded"]["frames_per_sec"], "replicated", g["replicated_map"]["frames_per_sec"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s5v_bench_n8.err").read()[-2500:])
PY
