# after turning whole-sector u_new stores on by default: new test, whole GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_copy_idle.py -m gpu -x -q > gpurun_out/final2_copyidle_test.log 2>&1; tail -n 15 gpurun_out/final2_copyidle_test.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final2_gpu_tests.log 2>&1; tail -n 4 gpurun_out/final2_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; tail -n 2 gpurun_out/final2_smoke.log
timeout 600 python bench.py > gpurun_out/final2_bench_n1.json 2> gpurun_out/final2_bench_n1.err; tail -c 600 gpurun_out/final2_bench_n1.err
python - <<'PY'
import json
for l in open('gpurun_out/final2_bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, d['clocks'])
        for w in d.get('other_workloads', []): print(' ', w.get('workload','')[:40], w.get('value',0)/1e9, w.get('roofline',{}).get('frac'))
PY
