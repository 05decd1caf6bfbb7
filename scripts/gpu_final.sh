# the round's final verification on one B200: full GPU suite, the default bench line, ncu evidence
# usage: gpu_final.sh TAG      then, here:  python scripts/update_profiles.py TAG gpurun_out/bench_TAG.json
TAG=${1:-r1}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/test_$TAG.log 2>&1
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
bash scripts/gpu_profile.sh $TAG > gpurun_out/profile_$TAG.log 2>&1
tail -3 gpurun_out/test_$TAG.log; tail -c 1500 gpurun_out/bench_$TAG.json
