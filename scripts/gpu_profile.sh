# ncu evidence for profiles/: launch list of the default bench command + full captures
mkdir -p gpurun_out
TAG=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 5 --propagate 0 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
for w in c2 c3 c4 c5; do
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 1 -o gpurun_out/prof_${w}_$TAG -f python bench.py --workload $w --steps 5 --warmup 5 --propagate 0 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full_$w.log 2>&1
done
ls -la gpurun_out | tail -8
