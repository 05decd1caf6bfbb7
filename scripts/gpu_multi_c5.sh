# usage: gpu_multi_c5.sh N [--test]  -- C5 weak-scaling bench line at N GPUs (optionally the 2-GPU IPC parity test first)
mkdir -p gpurun_out
N=$1
if [ "$2" = "--test" ]; then python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c5 --steps 100 --warmup 10 --no-extras $EXTRA 2> gpurun_out/bench_multi_c5_$N.err | grep '^{' > gpurun_out/bench_multi_c5_$N.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_multi_c5_$N.json"))
    print("c5 N=$N", round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"],3), "ms/step", d["e2e"] if "e2e" in d else "")
except Exception as e:
    print("c5 N=$N FAILED", e); print(open("gpurun_out/bench_multi_c5_$N.err").read()[-1500:])
PY
