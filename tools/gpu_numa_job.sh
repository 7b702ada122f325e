nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i "numa\|socket\|^CPU(s)" | head -8
for X in "" "--no-numa-bind"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 3 --no-gather $X > gpurun_out/fin4_bench.json 2> gpurun_out/fin4_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/fin4_bench.json").read().strip().splitlines()[-1])
    print("$X:", d["value"], round(d["ms_per_step"],4), "e2e", d["e2e"]["value"], round(d["e2e"]["ms_per_step"],3), d["e2e"]["h2d_gbs"], d["e2e"]["cpu_affinity"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/fin4_bench.err").read()[-2500:])
PY
done
