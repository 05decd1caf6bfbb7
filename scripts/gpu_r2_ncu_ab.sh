mkdir -p gpurun_out
FWB_NO_BRICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1010 -c 1 -o gpurun_out/prof_c5_nobrick -f python bench.py --workload c5 --steps 5 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full_c5_nobrick.log 2>&1
FWB_RING_BRICK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 505 -c 1 -o gpurun_out/prof_c3_nobrick -f python bench.py --workload c3 --steps 5 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full_c3_nobrick.log 2>&1
ls -la gpurun_out/prof_c5_nobrick.ncu-rep gpurun_out/prof_c3_nobrick.ncu-rep
