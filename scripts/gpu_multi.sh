# usage: gpu_multi.sh N [N2 ...]  -- 2-GPU IPC parity test, then weak-scaling bench lines
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
for N in "$@"; do
for w in c2 c5; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 100 --warmup 10 --no-extras > gpurun_out/bench_multi_${w}_$N.json 2> gpurun_out/bench_multi_${w}_$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_multi_${w}_$N.json"))
    print("$w N=$N", round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"],3), "ms/step", d["config"]["workload"][:60])
except Exception as e:
    print("$w N=$N FAILED", e); print(open("gpurun_out/bench_multi_${w}_$N.err").read()[-1500:])
PY
done
done
