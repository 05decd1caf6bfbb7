python -m pytest tests -m gpu -q -x -k "lr91 or luo or LuoRudy" 2>&1 | tail -4
python - <<'PY'
import sys; sys.path.insert(0, ".")
import torch, finitewave_b200 as fw
from finitewave_b200.devrun import DeviceSimulation
from finitewave_b200 import workloads
dev = torch.device("cuda")
for n, dim in ((4096, 2), (256, 3)):
    m = (fw.LuoRudy912D if dim == 2 else fw.LuoRudy913D)(); m.dt, m.dr, m.prog_bar = 0.01, 0.25, False
    shape = (n,) * dim
    sim = DeviceSimulation(m, workloads.fibrosis_mesh(shape, 0.0, 0, dev))
    st = fw.StimVoltageCoord2D(0, -20, 0, n, 0, 5) if dim == 2 else fw.StimVoltageCoord3D(0, -20, 0, n, 0, n, 0, 5)
    sim.add_stim(st)
    sim.run(300); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.run(50); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    K = 5 if dim == 2 else 7
    B = 129 + 8 * K
    print(f"LR91 {dim}D {n}^{dim} iso: {ms:.3f} ms/step -> {sim.n_myo/ms/1e6:.2f} G node-updates/s ({B*sim.n_myo/ms/1e6:.0f} GB/s algorithmic, {B*sim.n_myo/ms/1e6/6532.9:.3f} of HBM)")
PY
