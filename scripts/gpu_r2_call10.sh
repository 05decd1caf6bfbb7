mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cabi.py tests/test_gpu_parity.py tests/test_gpu_tiled.py tests/test_gpu_acceptance.py -m gpu -q 2>&1 | tail -4
timeout 120 python scripts/gpu_c1.py 2>&1 | tail -3
rm -f gpurun_out/variants.log
timeout 600 bash scripts/gpu_variants.sh "c5 c4 c5t lr91" base
STEPS=400 timeout 300 bash scripts/gpu_variants.sh "c2" base
STEPS=100 timeout 300 bash scripts/gpu_variants.sh "c3" base
