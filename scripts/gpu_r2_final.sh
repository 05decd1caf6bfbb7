# end-of-round verification on one B200: full GPU suite, smoke(), both bench arms as the driver
# runs them, ncu evidence from the propagated state
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1
tail -4 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -c 400 gpurun_out/final_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_bench.json").read().strip().split("\n")[-1])
r=json.loads(open("gpurun_out/final_ref.json").read().strip().split("\n")[-1])
print("value", round(d["value"]/1e9,3), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]/1e9,3), "ref", round(r["value"]/1e6,2), "M  same_config", d["config"]==r["config"], "wall", round(d["wall_s"],1))
for w in d["other_workloads"]:
    print("  ", w.get("workload","")[:70], "|", round(w.get("value",0)/1e9,3), round(w.get("hbm_frac",0),4), w.get("device_ms_per_1000_steps"), w.get("error"))
print(d["cpu_baseline"])
PY
# (ncu captures: scripts/gpu_profile_r2.sh TAG)
