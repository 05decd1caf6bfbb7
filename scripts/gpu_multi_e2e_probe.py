"""Where does the end-to-end leg of a multi-GPU run() spend its time?  (2 GPUs of one box)
1. raw host <-> device copies from pinned memory: one GPU at a time vs both at once
2. phases of TP063D.run(initialize=False) with model.devices = [0, 1]: upload, steps, download
"""
import importlib.util
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
n = torch.cuda.device_count()
print("devices", n, flush=True)
GB = 4
host = [torch.empty(GB * 2 ** 27, dtype=torch.float64, pin_memory=True) for _ in range(n)]
dev = [torch.empty(GB * 2 ** 27, dtype=torch.float64, device=f"cuda:{d}") for d in range(n)]
streams = [torch.cuda.Stream(device=d) for d in range(n)]


def sync():
    for d in range(n):
        torch.cuda.synchronize(d)


def copy(ds, h2d):
    sync()
    t0 = time.perf_counter()
    for d in ds:
        with torch.cuda.device(d), torch.cuda.stream(streams[d]):
            if h2d:
                dev[d].copy_(host[d], non_blocking=True)
            else:
                host[d].copy_(dev[d], non_blocking=True)
    sync()
    return len(ds) * GB * 2 ** 30 / (time.perf_counter() - t0) / 1e9


for h2d in (True, False):
    for ds in ([0], [1], [0, 1]) if n > 1 else ([0],):
        copy(ds, h2d)
        print("h2d" if h2d else "d2h", ds, "%.1f GB/s" % copy(ds, h2d), flush=True)
# both directions at once on different GPUs / on one GPU
if n > 1:
    sync(); t0 = time.perf_counter()
    with torch.cuda.device(0), torch.cuda.stream(streams[0]):
        dev[0].copy_(host[0], non_blocking=True)
    with torch.cuda.device(1), torch.cuda.stream(streams[1]):
        host[1].copy_(dev[1], non_blocking=True)
    sync(); print("h2d on 0 + d2h on 1: %.1f GB/s" % (2 * GB * 2 ** 30 / (time.perf_counter() - t0) / 1e9), flush=True)
del host, dev

spec = importlib.util.spec_from_file_location("fwb_bench", ROOT / "bench.py")
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
devices = list(range(min(n, 2)))
model = b.host_model("c5", 1.0, 32 * len(devices), devices if len(devices) > 1 else None)
model.t_max = 0.2
t0 = time.perf_counter(); model.run(); print("initialize + 20 steps %.2f s" % (time.perf_counter() - t0), flush=True)
import finitewave_b200.model as M
orig_up, orig_down = M.CardiacModel._upload, M.CardiacModel._download
acc = {"up": 0.0, "down": 0.0}


def up(self):
    t = time.perf_counter(); orig_up(self); sync(); acc["up"] += time.perf_counter() - t


def down(self, *a, **k):
    t = time.perf_counter(); orig_down(self, *a, **k); sync(); acc["down"] += time.perf_counter() - t


M.CardiacModel._upload, M.CardiacModel._download = up, down
for call in range(2):
    acc["up"] = acc["down"] = 0.0
    model.t_max = model.t + 100 * model.dt - 0.5 * model.dt
    sync(); t0 = time.perf_counter()
    model.run(initialize=False)
    sync(); tot = time.perf_counter() - t0
    nbytes = (2 + len(model._STATE)) * model.cardiac_tissue.mesh.size * 8
    print("call %d: total %.3f s, upload %.3f s (%.1f GB/s), download %.3f s (%.1f GB/s), rest %.3f s, %d steps"
          % (call, tot, acc["up"], nbytes / acc["up"] / 1e9, acc["down"], nbytes / acc["down"] / 1e9,
             tot - acc["up"] - acc["down"], model.gpu_steps), flush=True)
