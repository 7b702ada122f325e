# what the round-end driver does, in one gpurun call: full GPU parity suite, smoke, default bench line, reference arm
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/gpu_job.sh'
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1
tail -5 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -c 3000 gpurun_out/final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
tail -c 1200 gpurun_out/final_bench_ref.json
