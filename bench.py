#!/usr/bin/env python
"""
bench.py -- throughput of the fused per-time-step path (myocyte-node-updates/s, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3|c4|lr91]
                    [--impl b200|reference] [--no-extras]

A "step" is ONE explicit time step (one fused kernel launch: diffusion stencil + ionic ODE
+ trackers) over the whole tissue.  The default workload is the one BASELINE.json's metric
is quoted on -- C5, the TP06 3D anisotropic (19-point) slab cut along the fibre-rotation
axis: 128 x 1024 x 1024 nodes per GPU, i.e. 1024^3 (1.06e9 myocytes) at 8 GPUs (weak
scaling; one rank per GPU under torchrun, every rank owns one slab and exchanges one halo
plane per neighbour and step from inside the step kernel).  `--workload` selects another
BASELINE config.

One JSON line on stdout (rank 0):
  value      device-resident whole-job node-updates/s (CUDA events, max over ranks)
  e2e        the same metric through the public host API -- `TP063D.run()` on numpy arrays
             in pinned host memory: every timed call uploads u, u_new and all state arrays,
             runs `steps_per_call` time steps and downloads them again
  roofline   fused step kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json),
             algorithmic bytes per node from SURVEY.md 8d
  cpu_baseline  the REFERENCE's own numba path (oracle/_ref copy made by oracle/make_ref.py,
             `CardiacModel.run(initialize=False)`, all host threads) on a bounded sample of the
             same workload; the C/OpenMP oracle port's rate rides along as `port_value`
  other_workloads  C1 (launch-bound: steps/s), C2, C3, C4, C5 with the activation tracker, LR91
`--impl reference` prints the reference arm's line: the numba reference timed alone (the
oracle port only if the copy of the reference is missing, and then it says so).
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "myocyte-node-updates/sec (fp64, device-timed); % of HBM roofline"
UNIT = "node-updates/s"

# name -> (label, shape per GPU, stencil points, algorithmic bytes per node-update)
WORKLOADS = {
    "c5": ("C5 TP06 3D 1024x1024 x 128-per-GPU aniso 19-pt slab, cut along the fibre-rotation "
           "axis (1024^3 at 8 GPUs), face stimulus", [128, 1024, 1024], 19, 457),
    "c2": ("C2 Fenton-Karma 2D 4096x4096 aniso 9-pt, 30% random fibrosis", [4096, 4096], 9, 121),
    "c3": ("C3 Mitchell-Schaeffer 3D 512^3 iso 7-pt, focal stimulus, activation-time tracker "
           "every step", [512, 512, 512], 7, 97),
    "c4": ("C4 TP06 3D ventricle-shaped shell in 512^3, helix fibres, 19-pt", [512, 512, 512],
           19, 457),
    "lr91": ("LR91 2D 4096x4096 iso 5-pt, planar wave (extra, not a BASELINE config)",
             [4096, 4096], 5, 169),
    "c5t": ("C5 with an ActivationTime3DTracker sampling every step (TP06 + tracker)",
            [128, 1024, 1024], 19, 465),
}


def workload_config(name, world):
    """The `config` object: identical in the b200 arm and the reference arm."""
    label, shape, K, B = WORKLOADS[name]
    return {"workload": label, "shape_per_gpu": shape, "stencil_points": K,
            "bytes_per_node_update": B,
            "parallelism": "1 GPU" if world == 1 else
            f"{world} slabs along axis 0, 1-plane halo exchange per step"}


# ---------------------------------------------------------------------------
def measured_traffic(workload):
    """DRAM bytes per launch of the step kernel from the committed ncu capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        return float(json.loads(p.read_text())[workload]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock + throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------
# CPU side: the reference's numba path (and the oracle port) on a bounded sample
# ---------------------------------------------------------------------------
def port_time(workload, budget_s, steps=None, warmup=3):
    """The C/OpenMP oracle port (oracle/fw_oracle.c) on its bounded sample."""
    from oracle import oracle
    from oracle.bench_cases import cases_for_bench
    oracle.build()
    threads = os.cpu_count() or 1
    case, sample = cases_for_bench("c5" if workload == "c5t" else workload)
    if steps is None:
        sec, n_myo = oracle.time_steps(case, 3, n_threads=threads, warmup=1)
        per = sec / 3
        steps = int(min(2000, max(5, budget_s / max(per, 1e-6))))
    sec, n_myo = oracle.time_steps(case, steps, n_threads=threads, warmup=max(3, warmup))
    return dict(value=n_myo * steps / sec, unit=UNIT, cores=threads, kind="port",
                sample=f"{sample}, {steps} steps in {sec:.1f} s (oracle/fw_oracle.c, OpenMP)"), sec, steps


def reference_time(workload, steps, warmup):
    """The reference itself (numba) on its bounded sample -> cpu_baseline dict, sec, steps."""
    # torchrun sets OMP_NUM_THREADS=1 for its workers; the CPU arm is meant to use every core
    if os.environ.get("OMP_NUM_THREADS") in ("1", ""):
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import ref_numba
    r = ref_numba.time_reference("c5" if workload == "c5t" else workload, steps, warmup=warmup)
    return dict(value=r["n_myo"] * r["steps"] / r["seconds"], unit=UNIT, cores=r["threads"],
                kind="reference", sample=r["sample"],
                init_seconds=r["init_seconds"]), r["seconds"], r["steps"]


def cpu_baseline(workload, budget_s):
    """cpu_baseline object of the b200 arm's line: numba reference, port as a second field."""
    out = None
    try:
        # a first short run sizes the sample to ~budget_s of CPU work
        cb, sec, steps = reference_time(workload, 3, 2)
        per = sec / max(1, steps)
        more = int(min(1000, budget_s / max(per, 1e-6)))
        if more > 2 * steps:
            cb, sec, steps = reference_time(workload, more, 1)
        out = cb
    except Exception as e:
        out = {"error": "numba reference unavailable: " + repr(e)}
    try:
        pb, _, _ = port_time(workload, min(budget_s, 8.0))
        if "value" in out:
            out["port_value"], out["port_sample"] = pb["value"], pb["sample"]
        else:
            pb["note"] = out["error"]
            out = pb
    except Exception as e:
        out["port_error"] = repr(e)
    return out


def run_reference(args):
    """--impl reference: the reference's own numba CPU implementation of the path on the host
    cores (oracle/_ref; falls back to the oracle port if that copy is missing), W warm-up and
    K timed steps of a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    try:
        cb, sec, steps = reference_time(args.workload, args.steps, args.warmup)
    except Exception as e:
        cb, sec, steps = port_time(args.workload, 0, steps=args.steps, warmup=args.warmup)
        cb["note"] = "numba reference unavailable (" + repr(e) + "): oracle port timed instead"
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / max(1, steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, max(1, args.gpus)),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------
def time_device(sim, steps, warmup, dist=None, propagate=0):
    import torch
    if propagate > 0:
        sim.run(propagate)       # time from the state after >= 500 steps of propagation
    sim.run(warmup)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    sim.run(steps)
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    return ms, sim.launch_count() - l0


def halo_check(sim, rank, world, dist, device):
    """N > 1, after the timed steps: every slab's ghost slice must be, bit for bit, the owned
    boundary slice of the neighbour that stored it (the last in-kernel halo exchange has
    landed: DeviceSimulation.run ends with fwb_sim_halo_sync).  Each rank contributes an
    order-dependent fingerprint (sum and sum of squares of the slice, taken in the same
    element order on both sides) of its two ghost slices and its two owned boundary slices."""
    import torch
    u = sim.u_device()
    dist.barrier()

    def fp(x):
        x = x.reshape(-1)
        return [float(x.sum().item()), float((x * x).sum().item())]
    lo, hi = sim.halo
    mine = torch.tensor(
        (fp(u[0]) if lo else [0.0, 0.0]) + (fp(u[1 if lo else 0])) +
        (fp(u[-2 if hi else -1])) + (fp(u[-1]) if hi else [0.0, 0.0]),
        dtype=torch.float64, device=device)          # [ghost_lo, own_first, own_last, ghost_hi]
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    bad = 0
    for r in range(world - 1):
        a, b = allv[r].tolist(), allv[r + 1].tolist()
        if a[6:8] != b[2:4]:      # r's ghost_hi  vs  (r+1)'s first owned slice
            bad += 1
        if b[0:2] != a[4:6]:      # (r+1)'s ghost_lo  vs  r's last owned slice
            bad += 1
    return {"ok": bad == 0, "interfaces": world - 1, "mismatches": bad,
            "what": "ghost slice == neighbour's owned boundary slice after the timed steps "
                    "(bitwise fingerprints, all-gathered)"}


def host_model(workload, scale=1.0, slices=None, devices=None):
    """The workload through the PUBLIC host API (numpy tissue, model classes).  `slices`:
    thickness along axis 0 (C5 family: 128 per GPU); `devices`: model.devices."""
    import finitewave_b200 as fw

    def r32(x, lo=32):
        return max(lo, int(round(x * scale)) // 32 * 32)
    if workload == "c2":
        n = r32(4096, 64)
        rng = np.random.default_rng(2)
        tissue = fw.CardiacTissue2D([n, n])
        mesh = np.ones((n, n), dtype=np.int8)
        mesh[rng.random((n, n)) <= 0.30] = 2
        tissue.mesh = mesh
        f = np.empty((n, n, 2))
        f[..., 0], f[..., 1] = np.cos(0.25 * np.pi), np.sin(0.25 * np.pi)
        tissue.fibers = f
        model = fw.FentonKarma2D()
        seq = fw.StimSequence()
        seq.add_stim(fw.StimVoltageCoord2D(0, 1, 0, n, 0, 5))
    elif workload in ("c5", "c5t"):
        s, n = (slices or r32(128)), r32(1024)
        tissue = fw.CardiacTissue3D([s, n, n])
        phi = np.linspace(-np.pi / 3, np.pi / 2, s - 2)
        f = np.zeros((s, n, n, 3))
        f[1:-1, :, :, 1] = np.cos(phi)[:, None, None]
        f[1:-1, :, :, 2] = np.sin(phi)[:, None, None]
        tissue.fibers = f
        model = fw.TP063D()
        seq = fw.StimSequence()
        seq.add_stim(fw.StimVoltageCoord3D(0, -20, 0, s, 0, 5, 0, n))
        if workload == "c5t":
            tr = fw.ActivationTime3DTracker()
            tr.threshold, tr.step = -40, 1
            ts = fw.TrackerSequence()
            ts.add_tracker(tr)
            model.tracker_sequence = ts
    else:
        return None
    model.dt, model.dr, model.prog_bar = 0.01, 0.25, False
    model.cardiac_tissue = tissue
    model.stim_sequence = seq
    if devices is not None:
        model.devices = list(devices)
    return model


def e2e_host_api(workload, steps_per_call, calls, scale=1.0, slices=None, devices=None):
    """Public API with host buffers: CardiacModel.run() on numpy (pinned) arrays; with
    `devices` the one call drives a slab per GPU (finitewave_b200/multi.py)."""
    import torch
    model = host_model(workload, scale, slices, devices)
    if model is None:
        return None
    t_init = time.perf_counter()
    model.t_max = 0.2
    model.run()                                   # initialise + warm up (20 steps)
    init_s = time.perf_counter() - t_init
    n_myo = model._engine.n_myo
    n_arrays = 2 + len(model._STATE)              # u, u_new, every state array, each way
    per_call = n_arrays * int(np.prod(model.cardiac_tissue.mesh.shape)) * 8
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    launches = 0
    for _ in range(calls):
        model.t_max = model.t + steps_per_call * model.dt - 0.5 * model.dt
        model.run(initialize=False)               # upload -> steps -> download
        launches += model.gpu_launches
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    name = type(model).__name__
    if devices is not None:
        name += f" (model.devices = {list(devices)}: one process, one slab per GPU)"
        if not getattr(model._engine, "multi", False):
            raise RuntimeError("the slab-decomposed engine did not engage")
    return {"value": n_myo * steps_per_call * calls / sec, "unit": UNIT,
            "shape": list(model.cardiac_tissue.mesh.shape),
            "h2d_bytes_per_step": per_call, "d2h_bytes_per_step": per_call,
            "steps_per_call": steps_per_call, "calls": calls, "seconds": sec,
            "api": f"finitewave_b200.{name}.run(initialize=False) on pinned numpy arrays; one "
                   "e2e step = one run() call of steps_per_call time steps (bytes are per call)",
            "launches": launches, "initialize_seconds": init_s}


def e2e_multi_gpu(args, world):
    """N > 1, rank 0 only, after the other ranks have exited: the C5 family through
    `TP063D.run()` on `world` GPUs of this one process.  The slab thickness per GPU is the
    device-timed workload's (128) when the host can hold the pinned arrays (21 arrays of
    1024 x 1024 x 128 N doubles), else the largest power-of-two fraction that fits."""
    import psutil
    import torch
    # wait until the other ranks' contexts are gone
    for d in range(world):
        for _ in range(200):
            free, total = torch.cuda.mem_get_info(d)
            if free > 0.9 * total:
                break
            time.sleep(0.1)
    per_slice = 1024 * 1024 * 8 * 21
    avail = psutil.virtual_memory().available
    slices = int(round(128 * args.scale)) // 32 * 32 or 32
    # (also bounded to ~64 GB of pinned arrays so that the leg stays within a minute or two)
    while slices > 32 and (slices * world * per_slice * args.scale ** 2 * 1.35 > 0.6 * avail
                           or slices * world * per_slice * args.scale ** 2 > 64e9):
        slices //= 2
    out = e2e_host_api(args.workload, args.e2e_steps, 2, scale=args.scale, slices=slices * world,
                       devices=list(range(world)))
    out["slices_per_gpu"] = slices
    if slices != 128:
        out["note"] = (f"slab thickness reduced to {slices} slices per GPU for this leg (the "
                       "pinned host arrays of the whole tissue are bounded to ~64 GB / 60 % of RAM)")
    return out


def c1_line(device):
    """C1 (README quick start, AP 2D 100x100, 1000 steps): launch-latency-bound, 0.7 MB of
    state -- a roofline fraction is meaningless at this size, so it is reported as steps/s:
    device time of 1000 steps (CUDA events) and the wall time of the user's `model.run()`."""
    import torch
    import finitewave_b200 as fw

    def make():
        tissue = fw.CardiacTissue2D([100, 100])
        model = fw.AlievPanfilov2D()
        model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 10, False
        seq = fw.StimSequence()
        seq.add_stim(fw.StimVoltageCoord2D(0, 1, 1, 99, 1, 3))
        tr = fw.ActivationTime2DTracker()
        tr.threshold, tr.step = 0.5, 100
        ts = fw.TrackerSequence()
        ts.add_tracker(tr)
        model.cardiac_tissue, model.stim_sequence, model.tracker_sequence = tissue, seq, ts
        return model
    model = make()
    model.run()                                     # warm-up (allocations, graph capture)
    n_myo = model._engine.n_myo
    best_wall, best_dev = 1e9, 1e9
    for _ in range(5):
        model = make()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.run()                                 # initialize + 1000 steps + download
        best_wall = min(best_wall, time.perf_counter() - t0)
        # device time of 1000 more steps alone
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng = model._engine
        eng.set_time(model.t, model.step)
        e0.record()
        eng.run(1000)
        e1.record()
        torch.cuda.synchronize()
        best_dev = min(best_dev, e0.elapsed_time(e1) * 1e-3)
    return {"workload": "C1 Aliev-Panfilov 2D 100x100 iso 5-pt, planar wave, activation tracker "
                        "every 100 steps, 1000 steps (README quick start)",
            "steps": 1000, "myocyte_nodes": n_myo,
            "device_ms_per_1000_steps": best_dev * 1e3, "steps_per_s": 1000 / best_dev,
            "value": n_myo * 1000 / best_dev, "unit": UNIT,
            "e2e": {"value": n_myo * 1000 / best_wall, "unit": UNIT,
                    "seconds_per_run": best_wall,
                    "api": "finitewave_b200.AlievPanfilov2D.run() -- initialize(), upload, "
                           "1000 steps, download; best of 5"},
            "note": "launch-latency-bound: 9604 nodes, 0.7 MB; no roofline fraction at this size",
            "gpu_launches": int(model.gpu_launches)}


def run_b200(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: finitewave_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    real_stdout = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL's banner ("NCCL version ...") and debug
        # output are written to file descriptor 1 by the library itself, so descriptor 1 is
        # pointed at stderr for the duration of the run and the line goes to the saved one
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from finitewave_b200 import workloads

    t_wall0 = time.perf_counter()
    peak, peak_src = measured_peak()
    device = torch.device(f"cuda:{local}")
    sim, info = workloads.build(args.workload, device, scale=args.scale, rank=rank, world=world,
                                dist=dist)

    with ClockSampler(local) as clk:
        ms, launches = time_device(sim, args.steps, args.warmup, dist, args.propagate)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        nm = torch.tensor([info["n_myo"]], dtype=torch.float64, device=device)
        dist.all_reduce(nm, op=dist.ReduceOp.SUM)
        n_total = float(nm.item())
    else:
        ms_max, n_total = ms, float(info["n_myo"])
    halo = None
    if dist is not None:
        try:
            halo = halo_check(sim, rank, world, dist, device)
        except Exception as e:
            halo = {"ok": False, "error": repr(e)}
    value = n_total * args.steps / (ms_max * 1e-3)
    kernel_ms = ms / args.steps
    achieved = info["bytes_per_node"] * info["n_myo"] / (kernel_ms * 1e-3) / 1e9
    big = info["bytes_per_node"] * info["n_myo"] > 4 * 126e6
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, world),
        "workload_detail": {
            "built": info["workload"], "shape_per_gpu": info["shape"],
            "myocyte_nodes_per_gpu": info["n_myo"], "myocyte_nodes_total": int(n_total),
            "l2_policy": "working set per step >> 126 MB L2 (no flush needed)" if big else
            "working set fits L2; number is L2-resident",
            "state": f"after {args.propagate} propagation steps + {args.warmup} warm-up steps "
                     "from the stimulus at t = 0", "scale": args.scale},
        "gpu_launches": launches,
        "halo_check": halo,
        "clocks": clk.summary(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": measured_traffic(args.workload) if world == 1 and args.scale == 1.0 else None,
                     "algorithmic_bytes": info["bytes_per_node"] * info["n_myo"],
                     "peak_source": peak_src,
                     "kernel": "fwb::step_kernel (fused diffusion + ionic + trackers)",
                     "kernel_ms": kernel_ms},
    }
    del sim
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if world > 1:
        # the device-timed part is done: the other ranks leave, and rank 0 measures the
        # end-to-end number through the public API -- ONE process, `TP063D.run()` with
        # model.devices = all N GPUs (one slab per GPU, finitewave_b200/multi.py)
        dist.barrier()
        dist.destroy_process_group()
        dist = None
        if rank != 0:
            return
        print(f"[bench] device-timed part done ({value / 1e9:.2f} G upd/s); rank 0 continues alone",
              file=sys.stderr, flush=True)
        if not args.no_e2e:
            try:
                line["e2e"] = e2e_multi_gpu(args, world)
            except BaseException as e:          # never lose the device number
                line["e2e"] = {"error": repr(e)}
            print(f"[bench] e2e leg: {line['e2e']}", file=sys.stderr, flush=True)

    if rank == 0 and world == 1:
        import gc
        try:
            line["e2e"] = None if args.no_e2e else e2e_host_api(args.workload, args.e2e_steps, 2,
                                                                scale=args.scale)
        except Exception as e:                               # never lose the device number
            line["e2e"] = {"error": repr(e)}
        if line.get("e2e") is None:
            line["e2e"] = {"value": None, "unit": UNIT,
                           "note": "host-API e2e is measured for c5 and c2"}
        gc.collect()
        torch.cuda.empty_cache()
        if not args.no_extras:
            extras = []
            try:
                extras.append(c1_line(device))
            except Exception as e:
                extras.append({"workload": "c1", "error": repr(e)})
            for w in ("c5t", "c2", "c3", "c4", "c5", "lr91"):
                if w == args.workload:
                    continue
                try:
                    s2, i2 = workloads.build(w, device, scale=args.scale)
                    # a timed region of >= ~150 ms whatever the step costs (C2: 0.25 ms)
                    ms_probe, _ = time_device(s2, 10, 5, None, args.propagate)
                    k = int(max(10, min(1000, 150.0 / max(ms_probe / 10, 1e-3))))
                    ms2, l2 = time_device(s2, k, 5, None, 0)
                    ach = i2["bytes_per_node"] * i2["n_myo"] * k / (ms2 * 1e-3) / 1e9
                    extras.append({"workload": i2["workload"], "value": i2["n_myo"] * k / (ms2 * 1e-3),
                                   "unit": UNIT, "ms_per_step": ms2 / k, "steps": k,
                                   "myocyte_nodes": i2["n_myo"],
                                   "bytes_per_node_update": i2["bytes_per_node"],
                                   "hbm_gbs": ach, "hbm_frac": ach / peak, "gpu_launches": l2})
                    del s2
                except Exception as e:
                    extras.append({"workload": w, "error": repr(e)})
                gc.collect()
                torch.cuda.empty_cache()
            line["other_workloads"] = extras
        if not args.no_cpu:
            try:
                line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_budget)
            except Exception as e:
                line["cpu_baseline"] = {"error": repr(e)}
    line["wall_s"] = time.perf_counter() - t_wall0
    if rank == 0:
        if real_stdout is not None:
            sys.stdout.flush()
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=list(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every axis (debugging)")
    ap.add_argument("--e2e-steps", type=int, default=500,
                    help="time steps per model.run() call of the e2e leg")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--propagate", type=int, default=500,
                    help="untimed time steps before the warm-up (SURVEY 8d: >= 500)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
