"""Back-to-back steps of the ring kernel (C2, C3): kernel durations and the gaps between them from
CUPTI's activity records (torch.profiler sees every kernel of the process), against the CUDA-event
time of the same run."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from finitewave_b200 import workloads

dev = torch.device("cuda:0")
for w, steps in (("c2", 200), ("c3", 60)):
    sim, info = workloads.build(w, dev)
    sim.run(500)
    sim.run(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.run(steps); e1.record(); torch.cuda.synchronize()
    ev_us = e0.elapsed_time(e1) * 1e3 / steps
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        sim.run(steps)
        torch.cuda.synchronize()
    ks = sorted((e for e in prof.events() if "step_kernel" in e.name),
                key=lambda e: e.time_range.start)
    dur = [e.time_range.end - e.time_range.start for e in ks]
    gap = [b.time_range.start - a.time_range.end for a, b in zip(ks, ks[1:])]
    span = (ks[-1].time_range.end - ks[0].time_range.start) / len(ks)
    dur.sort(); gap.sort()
    print("%s: %d kernels; events %.1f us/step; CUPTI span %.1f us/step, kernel median %.1f us "
          "(min %.1f), gap median %.2f us (max %.1f)" %
          (w, len(ks), ev_us, span, dur[len(dur) // 2], dur[0], gap[len(gap) // 2], gap[-1]), flush=True)
    del sim
    torch.cuda.empty_cache()
