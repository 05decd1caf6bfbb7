# round 2, first GPU call: sanity tests, TP06 variant A/B, the new default bench line,
# the numba reference arm, one ncu capture of the TP06 kernel from the propagated state
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/call1_env.txt
python -c "import psutil,os; print('host cores', os.cpu_count(), 'ram GB', psutil.virtual_memory().total/2**30)" >> gpurun_out/call1_env.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/call1_tests.log
rm -f gpurun_out/variants.log
timeout 900 bash scripts/gpu_variants.sh "c5" base s1b3 s1b4 s1b4f s1b4fs s1b3s fast diet
timeout 300 bash scripts/gpu_variants.sh "c4" base s1b4f s1b3s
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/call1_ref.json 2> gpurun_out/call1_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/call1_bench.json 2> gpurun_out/call1_bench.err
tail -c 600 gpurun_out/call1_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 510 -c 1 -o gpurun_out/prof_c5_r2a -f python bench.py --workload c5 --steps 5 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_c5_r2a.log 2>&1
ls -la gpurun_out | tail -12
cat gpurun_out/variants.log
