# LR91 after sharing one root exponential between five slopes: parity + throughput
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "lr91 or luo or LuoRudy or heavy or tiled or fixtures or ionic" > gpurun_out/lr91_tests.log 2>&1; tail -n 3 gpurun_out/lr91_tests.log
timeout 300 python bench.py --workload lr91 --steps 100 --warmup 10 --no-e2e --no-cpu --no-extras > gpurun_out/lr91_bench.json 2> gpurun_out/lr91_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/lr91_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print('lr91', d['value']/1e9, d['roofline']['frac'], d['clocks'])
PY
