mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multidevice.py::test_slabs_on_different_gpus_match_single_gpu > gpurun_out/call8_tests.log 2>&1
tail -6 gpurun_out/call8_tests.log
rm -f gpurun_out/variants.log
timeout 300 bash scripts/gpu_variants.sh "c2 c3" base
echo "--- FWB_NO_BRICK=1" | tee -a gpurun_out/variants.log
FWB_NO_BRICK=1 timeout 300 bash scripts/gpu_variants.sh "c2 c3" base
