# session 5, run U (2 GPUs): snapshot wait deferred to the next light-volume write
python -m pytest tests/test_host_processors.py tests/test_sharding.py -m gpu -x -q 2>&1 | tail -2
for X in "--exchange auto" "--exchange auto --exchange-eager-wait" "--exchange nccl"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 3 --no-e2e --no-gather $X --check-exchange > gpurun_out/s5u_bench.json 2> gpurun_out/s5u_bench.err
echo "exit $?"
grep -h "exchange check\|unavailable" gpurun_out/s5u_bench.err | head -2
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5u_bench.json").read().strip().splitlines()[-1])
    print("$X:", d["value"], round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms_per_step"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/s5u_bench.err").read()[-2500:])
PY
done
