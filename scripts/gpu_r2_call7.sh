mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multidevice.py::test_slabs_on_different_gpus_match_single_gpu > gpurun_out/call7_tests.log 2>&1
tail -6 gpurun_out/call7_tests.log
rm -f gpurun_out/variants.log
timeout 300 bash scripts/gpu_variants.sh "c4" base
FWB_PACKED=0 timeout 300 bash scripts/gpu_variants.sh "c4" base
