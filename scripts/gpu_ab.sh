# usage: gpu_ab.sh "<flags1>" "<flags2>" ...  -- A/B of TP06 build flags (each a FWB_EXTRA_FLAGS
# value; "-" = none; a leading "VAR=val;" sets an environment variable for the bench) on C5
# and C4; only step_tp06.cu is rebuilt per variant
mkdir -p gpurun_out
for f in "$@"; do
  [ "$f" = "-" ] && f=""
  envs=""
  case "$f" in *";"*) envs="${f%%;*}"; f="${f#*;}";; esac
  rm -f finitewave_b200/_build/step_tp06.o
  FWB_EXTRA_FLAGS="$f" python -m finitewave_b200.build > /dev/null 2>&1 || echo BUILD FAILED
  echo "== flags: $f  env: $envs"
  for w in c5 c4; do
  env $envs python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']/1e9,3), 'G/s', round(d['ms_per_step'],3), 'ms')"
  done
done
rm -f finitewave_b200/_build/step_tp06.o
python -m finitewave_b200.build > /dev/null 2>&1
