TAG=${1:-s6i}
mkdir -p gpurun_out
V=abvariants
bash tools/gpu_ab_env.sh ${TAG} "A=1 --;CPM_B200_LIB=$V/t64/libcpm_b200.so CPM_HOST_LIB=$V/t64/libcpm_host.so --;CPM_B200_LIB=$V/t256/libcpm_b200.so CPM_HOST_LIB=$V/t256/libcpm_host.so --"
( python -m pytest tests/test_detector_splat.py tests/test_bound.py -m gpu -q --maxfail=10 ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
# memcheck over this session's new kernels (texture-bound tracer, wavefront kernels, copy_index_photons, PTX DDA loop)
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_bound.py::test_cuda_bounded_tracer_variants_bit_exact tests/test_bound.py::test_cuda_bounded_tracer_dense_and_empty_media tests/test_detector_splat.py -m gpu -q -x ) > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/${TAG}_memcheck.log
