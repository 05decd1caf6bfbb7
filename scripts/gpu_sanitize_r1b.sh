export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck2.log python -m pytest tests/test_gpu_parity.py tests/test_fibrosis_patterns.py tests/test_gpu_checkpoint.py tests/test_gpu_slab.py -m gpu -q -x -k "spiral_core or lat_period or sym or fibrosis or pattern or checkpoint or animation or tp06 or lr91 or slab" > gpurun_out/sanitizer_memcheck2.pytest.log 2>&1
tail -3 gpurun_out/sanitizer_memcheck2.pytest.log; tail -3 gpurun_out/sanitizer_memcheck2.log
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck2.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c5_tp06 or lr91_2d or tp06_3d_iso" > gpurun_out/sanitizer_racecheck2.pytest.log 2>&1
tail -2 gpurun_out/sanitizer_racecheck2.pytest.log; tail -3 gpurun_out/sanitizer_racecheck2.log
