"""C2: is the host launch loop ahead of the GPU?  Host-return time of sim.run(n) against the
device time, and where the long gaps between step kernels sit."""
import sys
import time
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from finitewave_b200 import workloads

dev = torch.device("cuda:0")
sim, info = workloads.build("c2", dev)
sim.run(520)
torch.cuda.synchronize()
for steps in (200, 200, 800):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); sim.run(steps); e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("steps %d: host returned after %.2f ms, all done after %.2f ms, events %.1f us/step"
          % (steps, t_host * 1e3, t_all * 1e3, e0.elapsed_time(e1) * 1e3 / steps), flush=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    sim.run(400)
    torch.cuda.synchronize()
ks = sorted((e for e in prof.events() if "step_kernel" in e.name), key=lambda e: e.time_range.start)
gaps = [(b.time_range.start - a.time_range.end, i) for i, (a, b) in enumerate(zip(ks, ks[1:]))]
print("long gaps (us, after kernel #):", [(round(g, 1), i) for g, i in gaps if g > 10][:20], flush=True)
others = sorted({e.name[:60] for e in prof.events() if "step_kernel" not in e.name})
print("other device activities:", others[:10])
