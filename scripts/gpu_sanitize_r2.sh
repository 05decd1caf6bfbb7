# compute-sanitizer memcheck over the round-2 kernels on small cases (tile kernel: brick, plain
# loads, packed, hand-over to the reference-statement kernel; ring kernel with the brick;
# multi-step cluster kernel; slabs in one process)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 420 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck.log python -m pytest "tests/test_gpu_tiled.py" "tests/test_gpu_multidevice.py::test_slabs_of_one_process_match_single_gpu" -m gpu -q -x > gpurun_out/r2_sanitizer_memcheck.pytest.log 2>&1
tail -3 gpurun_out/r2_sanitizer_memcheck.pytest.log; tail -4 gpurun_out/r2_sanitizer_memcheck.log
