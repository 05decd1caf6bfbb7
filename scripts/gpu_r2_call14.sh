mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
rm -f gpurun_out/variants.log
timeout 200 bash scripts/gpu_variants.sh "c5 c4" base
echo "--- FWB_NO_PERSIST=1" | tee -a gpurun_out/variants.log
FWB_NO_PERSIST=1 timeout 200 bash scripts/gpu_variants.sh "c5 c4" base
