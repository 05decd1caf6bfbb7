"""
oracle/ref_numba.py -- the REFERENCE ITSELF (numba CPU path) as a timed baseline.
TEST / MEASUREMENT INFRASTRUCTURE: imported by bench.py's `--impl reference` and
`cpu_baseline` legs only, never by finitewave_b200/.

Imports the unmodified reference package from /root/reference (build container) or from the
copy oracle/make_ref.py put under oracle/_ref/ (GPU box), with the visualisation-only modules
stubbed, builds a bounded sample of a BASELINE workload with the reference's
own classes and times `model.run(initialize=False)` (finitewave/core/model/cardiac_model.py:
130-189) with `numba.set_num_threads(os.cpu_count())`.
"""
import os
import sys
import time
import types
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_fw = None


def available():
    return Path("/root/reference/finitewave").is_dir() or (_HERE / "_ref" / "finitewave").is_dir()


def reference():
    """import finitewave (the reference), stubs first"""
    global _fw
    if _fw is not None:
        return _fw
    root = "/root/reference" if Path("/root/reference/finitewave").is_dir() else str(_HERE / "_ref")
    if not (Path(root) / "finitewave").is_dir():
        raise ImportError("no copy of the reference: run oracle/make_ref.py in the build container")
    for n in ["matplotlib", "matplotlib.pyplot", "natsort", "pyvista", "ffmpeg", "skimage",
              "skimage.measure"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["natsort"].natsorted = sorted
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    saved = sys.modules.get("finitewave")
    if saved is not None and "finitewave_b200" in getattr(saved, "__name__", ""):
        del sys.modules["finitewave"]      # the drop-in alias of the product, not the reference
    sys.path.insert(0, root)
    try:
        import finitewave as fw
    finally:
        sys.path.remove(root)
    if not str(Path(fw.__file__).resolve()).startswith(str(Path(root).resolve())):
        raise ImportError(f"`finitewave` resolved to {fw.__file__}, not the reference copy")
    _fw = fw
    return fw


def build_sample(workload):
    """-> (model, n_myocytes, description).  Same models / stencils / fibre fields / fibrosis
    rule as finitewave_b200/workloads.py, on a shape the host holds."""
    fw = reference()
    if workload == "c1":
        # README.rst:417-451, exactly
        tissue = fw.CardiacTissue2D([100, 100])
        model = fw.AlievPanfilov2D()
        model.dt, model.dr, model.t_max = 0.01, 0.25, 10
        stim = fw.StimVoltageCoord2D(0, 1, 1, 99, 1, 3)
        seq = fw.StimSequence()
        seq.add_stim(stim)
        tr = fw.ActivationTime2DTracker()
        tr.threshold, tr.step = 0.5, 100
        ts = fw.TrackerSequence()
        ts.add_tracker(tr)
        model.cardiac_tissue, model.stim_sequence, model.tracker_sequence = tissue, seq, ts
        desc = "C1 exactly (README quick start, 100x100, 1000 steps per run)"
    elif workload == "c2":
        n = 4096
        tissue = fw.CardiacTissue2D([n, n])
        rng = np.random.default_rng(2)
        mesh = np.ones((n, n), dtype=np.int8)
        mesh[rng.random((n, n)) <= 0.30] = 2
        tissue.mesh = mesh
        f = np.empty((n, n, 2))
        f[..., 0], f[..., 1] = np.cos(0.25 * np.pi), np.sin(0.25 * np.pi)
        tissue.fibers = f
        model = fw.FentonKarma2D()
        model.dt, model.dr = 0.01, 0.25
        seq = fw.StimSequence()
        seq.add_stim(fw.StimVoltageCoord2D(0, 1, 0, n, 0, 5))
        model.cardiac_tissue, model.stim_sequence = tissue, seq
        desc = f"C2 at full size {n}x{n} (FK aniso-9, 30% fibrosis)"
    elif workload == "c3":
        n = 256
        tissue = fw.CardiacTissue3D([n, n, n])
        model = fw.MitchellSchaeffer3D()
        model.dt, model.dr = 0.01, 0.25
        c = n // 2
        seq = fw.StimSequence()
        seq.add_stim(fw.StimVoltageCoord3D(0, 1, c - 5, c + 5, c - 5, c + 5, c - 5, c + 5))
        tr = fw.ActivationTime3DTracker()
        tr.threshold, tr.step = 0.5, 1
        ts = fw.TrackerSequence()
        ts.add_tracker(tr)
        model.cardiac_tissue, model.stim_sequence, model.tracker_sequence = tissue, seq, ts
        desc = f"C3 reduced to {n}^3 (MS iso-7, focal stimulus, activation-time tracker)"
    elif workload in ("c4", "c5"):
        # the C5 slab with its full thickness along the fibre-rotation axis, reduced laterally
        shape = [128, 160, 160]
        tissue = fw.CardiacTissue3D(shape)
        phi = np.linspace(-np.pi / 3, np.pi / 2, shape[0] - 2)
        f = np.zeros((*shape, 3))
        f[1:-1, :, :, 1] = np.cos(phi)[:, None, None]
        f[1:-1, :, :, 2] = np.sin(phi)[:, None, None]
        tissue.fibers = f
        model = fw.TP063D()
        model.dt, model.dr = 0.01, 0.25
        seq = fw.StimSequence()
        seq.add_stim(fw.StimVoltageCoord3D(0, -20, 0, shape[0], 0, 5, 0, shape[2]))
        model.cardiac_tissue, model.stim_sequence = tissue, seq
        desc = (f"C5 slab reduced laterally to {shape[0]}x{shape[1]}x{shape[2]} "
                "(TP06 aniso-19, same 128-slice thickness and fibre rotation)")
    else:
        raise ValueError(workload)
    model.prog_bar = False
    return model, desc


def time_reference(workload, steps, warmup=3, threads=None):
    """Times `steps` time steps of the reference's run() loop (after initialize() and
    `warmup` JIT-warming steps).  -> dict(seconds, n_myo, steps, threads, sample)"""
    import numba
    threads = int(threads or os.cpu_count() or 1)
    threads = min(threads, numba.config.NUMBA_NUM_THREADS)
    numba.set_num_threads(threads)
    model, desc = build_sample(workload)
    dt = model.dt
    t_init0 = time.perf_counter()
    model.t_max = max(1, warmup) * dt - 0.5 * dt
    model.run(num_of_theads=threads)                        # initialize + JIT + warm-up steps
    init_s = time.perf_counter() - t_init0
    n_myo = int(len(model.cardiac_tissue.myo_indexes))
    step0 = model.step
    model.t_max = model.t + steps * dt - 0.5 * dt
    t0 = time.perf_counter()
    model.run(initialize=False, num_of_theads=threads)
    sec = time.perf_counter() - t0
    done = model.step - step0
    return dict(seconds=sec, n_myo=n_myo, steps=int(done), threads=threads, init_seconds=init_s,
                sample=f"{desc}; {done} steps of the reference's CardiacModel.run(initialize=False) "
                       f"in {sec:.2f} s, numba {numba.__version__} on {threads} threads")
