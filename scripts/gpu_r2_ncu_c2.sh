# full ncu capture of the ring kernel on C2 with whole-sector u_new stores (the shipped default)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 505 -c 1 -o gpurun_out/prof_c2_r2d -f python bench.py --workload c2 --steps 5 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full_c2_r2d.log 2>&1
ls -la gpurun_out/prof_c2_r2d.ncu-rep
