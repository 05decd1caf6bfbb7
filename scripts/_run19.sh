python -m pytest tests -m gpu -q -x -s -k "tp06 or TP06 or device_math" 2>&1 | grep -v "^$" | tail -6
for w in c5 c4; do python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']/1e9,3), 'G/s', round(d['ms_per_step'],3), 'ms')"; done 2>&1 | tee gpurun_out/ab14.log
