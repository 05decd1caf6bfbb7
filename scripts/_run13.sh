mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/test_final_c.log 2>&1
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err
bash scripts/gpu_profile.sh r1_g > gpurun_out/profile_r1_g.log 2>&1
tail -3 gpurun_out/test_final_c.log; tail -c 1500 gpurun_out/bench_r1_g.json
