# session 5, run J: gather (2 whole records per iteration) parity + timing; launch list of our kernels; ncu full of the grid kernels and the gather
python -m pytest tests/test_gather.py tests/test_raycast.py tests/test_configs.py -m gpu -x -q 2>&1 | tail -2
for L in "--gather-interleaved" ""; do
python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu $L > gpurun_out/s5j_bench.json 2> gpurun_out/s5j_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5j_bench.json").read().strip().splitlines()[-1])
    g=d["gather"]; print("layout='$L'", round(g["photon_map_build_ms"],3), round(g["raymarch_ms"],3), round(g["frames_per_sec"],1), g["pixels_lit"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/s5j_bench.err").read()[-1500:])
PY
done
K='regex:(trace|onesweep|histogram|hist_scan|detect|splat|classify|minmax|minmax8|diff|diff8|range|range8|bound|tf_summary|select|reduce|seed_streams|fill_u32|scatter_fill|directional|mesh_intersect|uniform2d|cell_range|hash|gather|raycast|reorder|photon_cell_keys|mix)_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 800 --csv --log-file gpurun_out/s5j_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/s5j_launches.log 2>&1
python tools/launch_summary.py gpurun_out/s5j_launches.csv "python bench.py --steps 4 --warmup 3 --no-cpu   (-k <our kernels> -c 800; C4: resident leg = first frame + 3 warm-up + 4 timed frames, gather / final-image leg, e2e leg)" > gpurun_out/s5j_launches_summary.txt 2>&1; head -14 gpurun_out/s5j_launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:(minmax8|diff8|range8)_kernel' -c 6 -o gpurun_out/s5j_grids -f python tools/quickbench.py grids > gpurun_out/s5j_ncu_grids.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 1 -c 1 -o gpurun_out/s5j_gather -f python bench.py --steps 2 --warmup 1 --timesteps 6 --no-e2e --no-cpu --gather-interleaved > gpurun_out/s5j_ncu_gather.log 2>&1
ls gpurun_out
