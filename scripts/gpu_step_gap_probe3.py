"""C2: event-timed steps right after a long run vs after one second of idle (burst vs sustained),
with the SM clock NVML reports halfway through each."""
import sys
import threading
import time
from pathlib import Path

import pynvml
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from finitewave_b200 import workloads

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda:0")
sim, info = workloads.build("c2", dev)
sim.run(520)
torch.cuda.synchronize()


def timed(steps, label):
    clocks = []
    stop = threading.Event()

    def sample():
        while not stop.is_set():
            clocks.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                           pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.005)
    th = threading.Thread(target=sample)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    th.start()
    e0.record(); sim.run(steps); e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    c = sorted(x[0] for x in clocks); p = sorted(x[1] for x in clocks)
    print("%-28s %4d steps: %.1f us/step; SM clock median %d MHz (min %d), power median %.0f W (max %.0f)"
          % (label, steps, e0.elapsed_time(e1) * 1e3 / steps, c[len(c) // 2], c[0], p[len(p) // 2], p[-1]), flush=True)


timed(400, "after 520 steps, no pause")
timed(400, "again, no pause")
time.sleep(1.5)
timed(100, "after 1.5 s idle")
time.sleep(1.5)
timed(400, "after 1.5 s idle")
timed(2000, "long")
