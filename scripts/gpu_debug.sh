mkdir -p gpurun_out
python bench.py --workload c5 --scale 0.25 --steps 2 --warmup 3 --propagate 0 --no-e2e --no-cpu --no-extras > gpurun_out/dbg.out 2> gpurun_out/dbg.err
tail -c 1500 gpurun_out/dbg.err; tail -c 300 gpurun_out/dbg.out
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/dbg_san.log python bench.py --workload c5 --scale 0.125 --steps 1 --warmup 3 --propagate 0 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
head -60 gpurun_out/dbg_san.log
