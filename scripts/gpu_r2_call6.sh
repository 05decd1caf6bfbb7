mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_slab.py -m gpu -q 2>&1 | tail -5
