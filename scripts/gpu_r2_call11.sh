mkdir -p gpurun_out
timeout 120 python scripts/gpu_c1.py 2>&1 | tail -2 | cut -c150-330
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tiled.py tests/test_gpu_acceptance.py tests/test_gpu_fixtures.py tests/test_gpu_cabi.py tests/test_gpu_checkpoint.py -m gpu -q 2>&1 | tail -4
