# usage: gpu_prof_one.sh <workload> <tag>
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 1 -o gpurun_out/prof_$1_$2 -f python bench.py --workload $1 --steps 5 --warmup 5 --propagate 0 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full_$1.log 2>&1
ls -la gpurun_out/prof_$1_$2.ncu-rep
