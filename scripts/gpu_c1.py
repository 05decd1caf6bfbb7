"""C1 (README quick start): device time of 1000 steps and wall time of run(), vs the golden"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
line = bench.c1_line(torch.device("cuda"))
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in line.items() if k != "e2e"})
print(line["e2e"])
