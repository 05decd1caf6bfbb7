mkdir -p gpurun_out
for mb in 2 3 4; do
  FWB_EXTRA_FLAGS="-DFWB_TP06_MIN_BLOCKS=$mb" python -m finitewave_b200.build --force > /dev/null 2>&1
  grep -A2 "step_kernelINS_5ModelILi5EEELi3ELi1ELb0ELb0" finitewave_b200/_build/step_tp06.cu.ptxas.log | grep -E "registers|spill" | tr '\n' ' '
  echo "MIN_BLOCKS=$mb"
  python bench.py --workload c5 --steps 20 --warmup 5 --no-e2e --no-cpu --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['ms_per_step'])"
done
python -m finitewave_b200.build --force > /dev/null 2>&1
