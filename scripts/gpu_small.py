"""C1 (README quick start) and LR91 timing through the public API."""
import time, sys
sys.path.insert(0, ".")
import numpy as np, torch
import finitewave_b200 as fw

def c1():
    tissue = fw.CardiacTissue2D([100, 100])
    seq = fw.StimSequence(); seq.add_stim(fw.StimVoltageCoord2D(0, 1, 1, 99, 1, 3))
    ts = fw.TrackerSequence(); act = fw.ActivationTime2DTracker(); act.threshold = 0.5; act.step = 100
    ts.add_tracker(act)
    m = fw.AlievPanfilov2D(); m.dt, m.dr, m.t_max, m.prog_bar = 0.01, 0.25, 10, False
    m.cardiac_tissue, m.stim_sequence, m.tracker_sequence = tissue, seq, ts
    m.run()
    for _ in range(3):
        t0 = time.perf_counter(); m.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"C1 run() 1000 steps: {dt*1e3:.2f} ms  -> {9604*1000/dt/1e6:.1f} M node-updates/s, u.sum={m.u.sum():.9f}")
    eng = m._engine
    eng.set_time(0, 0)
    torch.cuda.synchronize(); t0 = time.perf_counter(); eng.run(10000); torch.cuda.synchronize()
    print(f"device loop only: {(time.perf_counter()-t0)/10000*1e6:.2f} us/step")

def lr91(n=2048):
    from finitewave_b200.devrun import DeviceSimulation
    from finitewave_b200 import workloads
    dev = torch.device("cuda")
    m = fw.LuoRudy912D(); m.dt, m.dr, m.prog_bar = 0.01, 0.25, False
    sim = DeviceSimulation(m, workloads.fibrosis_mesh((n, n), 0.0, 0, dev))
    sim.add_stim(fw.StimVoltageCoord2D(0, -20, 0, n, 0, 5))
    sim.run(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.run(50); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"LR91 2D {n}^2 iso: {ms:.3f} ms/step -> {sim.n_myo/ms/1e6:.2f} G node-updates/s ({169*sim.n_myo/ms/1e6:.0f} GB/s algorithmic)")

c1(); lr91()
