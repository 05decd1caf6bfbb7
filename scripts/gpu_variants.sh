# usage: gpu_variants.sh <workloads> <variant> ...   ("base" = the default library)
# one bench line per (variant, workload): device-resident value only
mkdir -p gpurun_out
WL=$1; shift
for v in "$@"; do
  lib=finitewave_b200/libfinitewave_b200_$v.so
  [ "$v" = "base" ] && lib=finitewave_b200/libfinitewave_b200.so
  for w in $WL; do
    FWB_LIB=$lib python bench.py --workload $w --steps ${STEPS:-30} --warmup 5 --no-e2e --no-cpu --no-extras 2>gpurun_out/var_${v}_${w}.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', '$w', round(d['value']/1e9,3), 'G/s', round(d['ms_per_step'],3), 'ms', 'frac', round(d['roofline']['frac'],4), d['clocks'])" | tee -a gpurun_out/variants.log
  done
done
