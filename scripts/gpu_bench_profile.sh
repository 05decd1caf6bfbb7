mkdir -p gpurun_out
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err
tail -c 3000 gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_a.csv python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 2 -o gpurun_out/prof_c2_r1_a python bench.py --steps 10 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 2 -o gpurun_out/prof_c5_r1_a python bench.py --workload c5 --steps 5 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full5.log 2>&1
ls -la gpurun_out
