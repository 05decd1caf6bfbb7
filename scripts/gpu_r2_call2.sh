# round 2, second GPU call: the compact-lane tile kernel (bricks by tensor TMA) -- parity suite,
# A/B of brick vs plain loads and 3 vs 4 blocks per SM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/call2_tests.log
rm -f gpurun_out/variants.log
timeout 600 bash scripts/gpu_variants.sh "c5 c4" base mb4 nosmem
echo "--- FWB_NO_BRICK=1" | tee -a gpurun_out/variants.log
FWB_NO_BRICK=1 timeout 600 bash scripts/gpu_variants.sh "c5 c4" base mb4
timeout 300 python bench.py --workload lr91 --steps 30 --warmup 5 --no-e2e --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lr91', round(d['value']/1e9,3), 'G/s frac', round(d['roofline']['frac'],4))" | tee -a gpurun_out/variants.log
timeout 300 python bench.py --workload c5t --steps 30 --warmup 5 --no-e2e --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5t', round(d['value']/1e9,3), 'G/s frac', round(d['roofline']['frac'],4))" | tee -a gpurun_out/variants.log
cat gpurun_out/variants.log
