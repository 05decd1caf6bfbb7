mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_multidevice.py -m gpu -q --timeout=90 > gpurun_out/call5_tests.log 2>&1
grep -n "^E  \|Error\|passed\|failed\|Timeout" gpurun_out/call5_tests.log | head -40
