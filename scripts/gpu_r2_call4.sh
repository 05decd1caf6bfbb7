mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fixtures.py tests/test_gpu_fullsize.py tests/test_gpu_tiled.py tests/test_gpu_cabi.py -m gpu -q > gpurun_out/call4_tests.log 2>&1
tail -8 gpurun_out/call4_tests.log
rm -f gpurun_out/variants.log
timeout 600 bash scripts/gpu_variants.sh "c5" base nobar
