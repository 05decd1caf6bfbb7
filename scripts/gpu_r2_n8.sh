mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 30 --warmup 5 --e2e-steps 200 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "exit code $?"
grep "^\[bench\]" gpurun_out/n${N}_bench.err | cut -c1-700
python - <<PY
import json
d=json.loads(open("gpurun_out/n${N}_bench.json").read().strip().split("\n")[-1])
print("N=$N value", round(d["value"]/1e9,3), "G/s ms/step", round(d["ms_per_step"],3), "nodes", d["workload_detail"]["myocyte_nodes_total"], "clocks", d["clocks"])
PY
free -g | head -2
