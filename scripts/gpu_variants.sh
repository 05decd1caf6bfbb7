# usage: gpu_variants.sh "<flags1>" "<flags2>" ...   (each a FWB_EXTRA_FLAGS value; "-" = none)
mkdir -p gpurun_out
for f in "$@"; do
  [ "$f" = "-" ] && f=""
  FWB_EXTRA_FLAGS="$f" python -m finitewave_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
  echo "== flags: $f"
  for w in c5 c4; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']/1e9,3), 'G/s', round(d['ms_per_step'],3), 'ms')"
  done
done
python -m finitewave_b200.build --force > /dev/null 2>&1
