mkdir -p gpurun_out
rm -f gpurun_out/variants.log
for rep in 1 2; do
echo "--- default (ring brick 3D only)" | tee -a gpurun_out/variants.log
STEPS=400 timeout 300 bash scripts/gpu_variants.sh "c2" base
STEPS=100 timeout 300 bash scripts/gpu_variants.sh "c3" base
echo "--- FWB_RING_BRICK=0" | tee -a gpurun_out/variants.log
FWB_RING_BRICK=0 STEPS=100 timeout 300 bash scripts/gpu_variants.sh "c3" base
echo "--- FWB_RING_BRICK=1" | tee -a gpurun_out/variants.log
FWB_RING_BRICK=1 STEPS=400 timeout 300 bash scripts/gpu_variants.sh "c2" base
done
timeout 400 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_cabi.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -3
