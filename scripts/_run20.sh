python -m pytest tests -m gpu -q -x -k "tp06 or TP06 or device_math or lr91 or luo or court" 2>&1 | tail -4
for w in c5 c4; do python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']/1e9,3), 'G/s', round(d['ms_per_step'],3), 'ms')"; done 2>&1 | tee gpurun_out/ab15.log
