# session 5, run P: full GPU parity suite, smoke, default bench line, reference arm
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s5p_pytest.log 2>&1
tail -5 gpurun_out/s5p_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s5p_smoke.log 2>&1; tail -2 gpurun_out/s5p_smoke.log
python bench.py > gpurun_out/s5p_bench.json 2> gpurun_out/s5p_bench.err
tail -c 5000 gpurun_out/s5p_bench.json; tail -3 gpurun_out/s5p_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s5p_bench_ref.json 2> gpurun_out/s5p_bench_ref.err
tail -c 1500 gpurun_out/s5p_bench_ref.json
