mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 5 --e2e-steps 200 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "exit code $?"
grep -v "^$" gpurun_out/n2_bench.err | tail -12 | cut -c1-600
wc -c gpurun_out/n2_bench.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/n2_bench.json"))
print("N=2 value", round(d["value"]/1e9,3), "G/s ms/step", round(d["ms_per_step"],3), "e2e", d["e2e"])
PY
free -g | head -2
