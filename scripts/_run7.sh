bash scripts/gpu_ab.sh - "-DFWB_NO_STAGE" "-DFWB_TP06_MIN_BLOCKS=3" 2>&1 | tee gpurun_out/ab5.log
for w in c2 c3; do python bench.py --workload $w --steps 50 --warmup 10 --no-e2e --no-cpu --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']/1e9,3), 'G/s', round(d['ms_per_step'],3), 'ms', d['roofline']['frac'])"; done 2>&1 | tee -a gpurun_out/ab5.log
python -m pytest tests -m gpu -q -x -k "tp06 or TP06 or lr91 or luo or court or slab or cabi" 2>&1 | tail -5 | tee gpurun_out/test6.log
