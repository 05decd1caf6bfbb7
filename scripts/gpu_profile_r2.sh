# ncu evidence for profiles/ (round 2): launch list of the default bench command + one full
# capture per workload of the step kernel FROM THE PROPAGATED STATE (500 steps + 5 warm-up)
# usage: gpu_profile_r2.sh TAG
mkdir -p gpurun_out
TAG=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 5 --propagate 100 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_launch_$TAG.log 2>&1
# tile-kernel workloads launch two kernels per step (fast + reference-statement): skip 2 * 505
for w in c5 c4; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1010 -c 1 -o gpurun_out/prof_${w}_$TAG -f python bench.py --workload $w --steps 5 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full_${w}_$TAG.log 2>&1
done
for w in c2 c3; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 505 -c 1 -o gpurun_out/prof_${w}_$TAG -f python bench.py --workload $w --steps 5 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full_${w}_$TAG.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -6
