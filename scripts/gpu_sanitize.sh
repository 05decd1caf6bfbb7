mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for tool in memcheck racecheck; do
timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log python -m pytest tests/test_gpu_slab.py tests/test_gpu_cabi.py "tests/test_gpu_parity.py::test_cuda_matches_oracle" -m gpu -q -x > gpurun_out/sanitizer_$tool.pytest.log 2>&1
tail -2 gpurun_out/sanitizer_$tool.pytest.log; tail -4 gpurun_out/sanitizer_$tool.log
done
