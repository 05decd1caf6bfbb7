mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/test2.log; tail -15 gpurun_out/test2.log
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err
tail -c 3500 gpurun_out/bench_r1_b.json; tail -5 gpurun_out/bench_r1_b.err
