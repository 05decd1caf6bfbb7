mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multidevice.py::test_slabs_on_different_gpus_match_single_gpu 2>&1 | tail -4
rm -f gpurun_out/variants.log
STEPS=400 timeout 300 bash scripts/gpu_variants.sh "c2" base
FWB_NO_COPY_IDLE=1 STEPS=400 timeout 300 bash scripts/gpu_variants.sh "c2" base
