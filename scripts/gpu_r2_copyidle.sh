# C2 with and without whole-sector u_new stores (FWB_COPY_IDLE=1): timing, DRAM bytes, parity
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum
for v in 0 1; do
  FWB_COPY_IDLE=$v timeout 300 python bench.py --workload c2 --steps 400 --warmup 20 --no-e2e --no-cpu --no-extras > gpurun_out/copyidle_c2_$v.json 2> gpurun_out/copyidle_c2_$v.err
  FWB_COPY_IDLE=$v timeout 300 ncu --metrics $M --clock-control none -k regex:step_kernel -s 505 -c 1 --csv --log-file gpurun_out/copyidle_ncu_$v.csv python bench.py --workload c2 --steps 5 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/copyidle_ncu_$v.log 2>&1
done
FWB_COPY_IDLE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tiled.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/copyidle_tests.log 2>&1
tail -n 3 gpurun_out/copyidle_tests.log
for v in 0 1; do python - <<PY
import json
for l in open('gpurun_out/copyidle_c2_$v.json'):
    if l.startswith('{'):
        d=json.loads(l); print('copy_idle=$v', d['value']/1e9, d['roofline']['frac'], d['clocks'])
PY
grep -v "^==" gpurun_out/copyidle_ncu_$v.csv | tail -n 6 | cut -d, -f5,13-
done
