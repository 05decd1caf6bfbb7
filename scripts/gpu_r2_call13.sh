mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multidevice.py::test_slabs_on_different_gpus_match_single_gpu 2>&1 | tail -3
rm -f gpurun_out/variants.log
timeout 300 bash scripts/gpu_variants.sh "c5 c4" base
