import sys; sys.path.insert(0, ".")
import torch, finitewave_b200 as fw
from finitewave_b200.devrun import DeviceSimulation
from finitewave_b200 import workloads
dev = torch.device("cuda")
n = 3072
m = fw.Courtemanche2D(); m.dt, m.dr, m.prog_bar = 0.01, 0.25, False
sim = DeviceSimulation(m, workloads.fibrosis_mesh((n, n), 0.0, 0, dev))
sim.add_stim(fw.StimVoltageCoord2D(0, -20, 0, n, 0, 5))
sim.run(200); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sim.run(40); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 40
B = 8 + 8 + 1 + 8 * 5 + 8 * 2 * 21
print(f"Courtemanche 2D {n}^2 iso: {ms:.3f} ms/step -> {sim.n_myo/ms/1e6:.2f} G/s ({B*sim.n_myo/ms/1e6/6532.9:.3f} of HBM, B={B})")
