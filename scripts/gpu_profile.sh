mkdir -p gpurun_out
python -m pytest tests/test_gpu_slab.py -m gpu -q -x 2>&1 | tail -5
for w in c2 c3 c5; do
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 1 -o gpurun_out/prof_${w}_r1_b -f python bench.py --workload $w --steps 5 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full_$w.log 2>&1
done
ls -la gpurun_out | tail -8
