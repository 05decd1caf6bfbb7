#!/usr/bin/env python
"""
bench.py -- throughput of the fused per-time-step path (myocyte-node-updates/s, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5]
                    [--impl b200|reference] [--no-extras]

A "step" is ONE explicit time step (one fused kernel launch: diffusion stencil +
ionic ODE + trackers) over the whole tissue.  The N=1 workload is BASELINE.json's
configs[1] (C2: Fenton-Karma 2D 4096x4096, anisotropic 9-point stencil, 30 % random
fibrosis); `--workload` selects another BASELINE config.  With N > 1 (torchrun, one
rank per GPU) every rank owns a slab of an N-times larger tissue (weak scaling) and
exchanges one halo plane per neighbour per step.

One JSON line on stdout (rank 0):
  value      device-resident whole-job node-updates/s (CUDA events, max over ranks)
  e2e        the same metric through the public host API (model.run() on numpy
             arrays in pinned host memory): every timed call uploads u, u_new and
             all state arrays, runs `steps_per_call` time steps (1000, the run length
             of the reference's README quick start), downloads them again
  roofline   fused step kernel vs the measured HBM copy bandwidth
             (MEASURED_PEAKS.json), algorithmic bytes per node from SURVEY.md 8d
  cpu_baseline  the CPU oracle port (oracle/, OpenMP, all host threads) timed on a
             bounded sample of the same workload (N=1, rank 0 only)
`--impl reference` times that CPU port alone (the reference is Python + numba and
cannot be built or shipped to the GPU box).
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
# CPU side: the oracle port on a bounded sample of the workload
# ---------------------------------------------------------------------------
def cpu_case(workload):
    """Reduced-shape twin of a workload for the CPU legs."""
    from oracle.bench_cases import cases_for_bench
    return cases_for_bench(workload)


def cpu_time(workload, budget_s, steps=None, warmup=3):
    from oracle import oracle
    oracle.build()
    threads = os.cpu_count() or 1
    case, sample = cpu_case(workload)
    if steps is None:
        sec, n_myo = oracle.time_steps(case, 3, n_threads=threads, warmup=1)
        per = sec / 3
        steps = int(min(2000, max(5, budget_s / max(per, 1e-6))))
    sec, n_myo = oracle.time_steps(case, steps, n_threads=threads, warmup=max(3, warmup))
    return dict(value=n_myo * steps / sec, unit=UNIT, cores=threads, kind="port",
                sample=f"{sample}, {steps} steps in {sec:.1f} s (oracle/fw_oracle.c, OpenMP)"), sec, steps


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the numba
    reference cannot travel) on the host cores, K timed steps of a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb, sec, steps = cpu_time(args.workload, 0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.workload], "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


WORKLOAD_NAMES = {
    "c2": "C2 Fenton-Karma 2D 4096x4096 aniso 9-pt, 30% random fibrosis",
    "c3": "C3 Mitchell-Schaeffer 3D 512^3 iso 7-pt, focal stimulus, activation-time tracker",
    "c4": "C4 TP06 3D ventricle-shaped shell in 512^3, helix fibres, 19-pt",
    "c5": "C5 TP06 3D 1024x1024x128-per-GPU aniso 19-pt slab",
}


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


def e2e_host_api(workload, steps_per_call, calls, scale=1.0):
    """Public API with host buffers: CardiacModel.run() on numpy (pinned) arrays."""
    import torch
    import finitewave_b200 as fw
    if workload != "c2":
        return None
    n = max(64, int(round(4096 * scale)) // 32 * 32)
    rng = np.random.default_rng(2)
    tissue = fw.CardiacTissue2D([n, n])
    mesh = np.ones((n, n), dtype=np.int8)
    mesh[rng.random((n, n)) <= 0.30] = 2
    tissue.mesh = mesh
    f = np.empty((n, n, 2))
    f[..., 0], f[..., 1] = np.cos(0.25 * np.pi), np.sin(0.25 * np.pi)
    tissue.fibers = f
    model = fw.FentonKarma2D()
    model.dt, model.dr, model.prog_bar = 0.01, 0.25, False
    model.cardiac_tissue = tissue
    seq = fw.StimSequence()
    seq.add_stim(fw.StimVoltageCoord2D(0, 1, 0, n, 0, 5))
    model.stim_sequence = seq
    model.t_max = 0.2
    model.run()                                   # initialise + warm up (20 steps)
    n_myo = model._engine.n_myo
    per_call = 4 * n * n * 8                      # u, u_new, v, w each way
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    launches = 0
    for _ in range(calls):
        model.t_max = model.t + steps_per_call * model.dt - 0.5 * model.dt
        model.run(initialize=False)               # upload -> steps -> download
        launches += model.gpu_launches
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    return {"value": n_myo * steps_per_call * calls / sec, "unit": UNIT,
            "h2d_bytes_per_step": per_call, "d2h_bytes_per_step": per_call,
            "steps_per_call": steps_per_call, "calls": calls,
            "api": "finitewave_b200.FentonKarma2D.run(initialize=False) on pinned numpy arrays; "
                   "one e2e step = one run() call of steps_per_call time steps",
            "launches": launches}


def e2e_slabs(sim, info, steps_per_call, calls, dist, device):
    """N > 1: every rank round-trips its slab (u, u_new, all state arrays) through pinned
    host memory around each run of `steps_per_call` steps; max over ranks."""
    import torch
    bufs = sim.host_buffers()
    sim.download_host(bufs)
    per_call = sum(b.numel() * 8 for b in bufs.values())
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(calls):
        sim.upload_host(bufs)
        sim.run(steps_per_call)
        sim.download_host(bufs)
    torch.cuda.synchronize()
    sec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    nm = torch.tensor([info["n_myo"]], dtype=torch.float64, device=device)
    dist.all_reduce(nm, op=dist.ReduceOp.SUM)
    return {"value": float(nm.item()) * steps_per_call * calls / float(sec.item()), "unit": UNIT,
            "h2d_bytes_per_step": per_call, "d2h_bytes_per_step": per_call,
            "steps_per_call": steps_per_call, "calls": calls,
            "api": "finitewave_b200.devrun.DeviceSimulation upload_host -> run -> download_host "
                   "per rank (bytes are per rank and call); one e2e step = one call"}


def run_b200(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: finitewave_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: keep NCCL's banner / debug output off it
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
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
    value = n_total * args.steps / (ms_max * 1e-3)
    kernel_ms = ms / args.steps
    achieved = info["bytes_per_node"] * info["n_myo"] / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": info["workload"], "shape_per_gpu": info["shape"],
                   "myocyte_nodes_per_gpu": info["n_myo"], "stencil_points": info["K"],
                   "bytes_per_node_update": info["bytes_per_node"],
                   "l2_policy": "working set per step >> 126 MB L2 (no flush needed)"
                   if info["bytes_per_node"] * info["n_myo"] > 4 * 126e6 else
                   "working set fits L2; number is L2-resident",
                   "state": f"after {args.propagate} propagation steps + {args.warmup} warm-up steps "
                            "from the stimulus at t = 0",
                   "parallelism": "1 GPU" if world == 1 else f"{world} slabs, 1-plane halo exchange"},
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": measured_traffic(args.workload) if world == 1 and args.scale == 1.0 else None,
                     "algorithmic_bytes": info["bytes_per_node"] * info["n_myo"],
                     "peak_source": peak_src,
                     "kernel": "fwb::step_kernel (fused diffusion + ionic + trackers)",
                     "kernel_ms": kernel_ms},
    }
    if world > 1 and not args.no_e2e:
        try:
            e2e = e2e_slabs(sim, info, args.e2e_steps, 2, dist, device)
        except Exception as e:
            e2e = {"error": repr(e)}
        line["e2e"] = e2e
    del sim
    torch.cuda.empty_cache()

    if rank == 0 and world == 1:
        try:
            line["e2e"] = None if args.no_e2e else e2e_host_api(args.workload, args.e2e_steps, 2,
                                                                scale=args.scale)
        except Exception as e:                               # never lose the device number
            line["e2e"] = {"error": repr(e)}
        if line.get("e2e") is None:
            line["e2e"] = {"value": None, "unit": UNIT, "note": "host-API e2e is measured on C2"}
        torch.cuda.empty_cache()
        if not args.no_extras:
            extras = []
            for w in ("c3", "c4", "c5", "lr91"):
                if w == args.workload:
                    continue
                try:
                    s2, i2 = workloads.build(w, device, scale=args.scale)
                    ms2, l2 = time_device(s2, max(10, args.steps // 8), 5, None, args.propagate)
                    k = max(10, args.steps // 8)
                    ach = i2["bytes_per_node"] * i2["n_myo"] * k / (ms2 * 1e-3) / 1e9
                    extras.append({"workload": i2["workload"], "value": i2["n_myo"] * k / (ms2 * 1e-3),
                                   "unit": UNIT, "ms_per_step": ms2 / k, "steps": k,
                                   "myocyte_nodes": i2["n_myo"],
                                   "bytes_per_node_update": i2["bytes_per_node"],
                                   "hbm_gbs": ach, "hbm_frac": ach / peak, "gpu_launches": l2})
                    del s2
                except Exception as e:
                    extras.append({"workload": w, "error": repr(e)})
                torch.cuda.empty_cache()
            line["other_workloads"] = extras
        if not args.no_cpu:
            try:
                cb, _, _ = cpu_time(args.workload, args.cpu_budget)
                line["cpu_baseline"] = cb
            except Exception as e:
                line["cpu_baseline"] = {"error": repr(e)}
    line["wall_s"] = time.perf_counter() - t_wall0
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOAD_NAMES))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every axis (debugging)")
    ap.add_argument("--e2e-steps", type=int, default=1000,
                    help="time steps per model.run() call of the e2e leg (README quick start: 1000)")
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
