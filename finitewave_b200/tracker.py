"""
Trackers.  Same names / attributes / gate as the reference
(finitewave/core/tracker/tracker.py:7-101, tracker_sequence.py;
ActivationTime  cpuwave2D/tracker/activation_time_2d_tracker.py:7-75,
ECG             cpuwave2D/tracker/ecg_2d_tracker.py:9-152, cpuwave3D/tracker/ecg_3d_tracker.py:9-140,
point samplers  cpuwave2D/tracker/{action_potential,multi_variable,variable}_2d_tracker.py).

Native trackers (``_native = True``) are evaluated on the device inside the step
runner -- activation time and ECG are fused into the step kernel -- and only
their small outputs come back.  Device-hook trackers (``_device_hook = True``:
LocalActivationTime, Period) run their own small kernels on the device potential at
their sample steps; nothing is downloaded until their output is read.  Any other
Tracker subclass is a host hook: the model downloads ``state_vars`` before calling
its ``_track()``.
"""
import copy
import ctypes
from pathlib import Path

import numpy as np
import torch

from ._lib import check, shape_arr


class Tracker:
    _native = False
    _device_hook = False

    def __init__(self):
        self.model = None
        self.file_name = "tracked_data"
        self.path = "."
        self.start_time = 0
        self.end_time = np.inf
        self.step = 1

    def initialize(self, model):
        raise NotImplementedError

    def _track(self):
        raise NotImplementedError

    def gate(self, t, step):
        """Tracker.track's condition (tracker.py:70-84)."""
        if self.start_time > t or t > self.end_time:
            return False
        return step % self.step == 0

    def track(self):
        if self.gate(self.model.t, self.model.step):
            self._track()

    def clone(self):
        return copy.deepcopy(self)

    def write(self):
        np.save(Path(self.path, self.file_name).with_suffix(".npy"), self.output)

    # native trackers override
    def _register(self, engine, model, max_samples):
        raise NotImplementedError

    def _collect(self, engine):
        pass


class TrackerSequence:
    def __init__(self):
        self.sequence = []
        self.model = None

    def initialize(self, model):
        self.model = model
        for tracker in self.sequence:
            tracker.initialize(model)

    def add_tracker(self, tracker):
        self.sequence.append(tracker)

    def remove_trackers(self):
        self.sequence = []

    def tracker_next(self):
        for tracker in self.sequence:
            tracker.track()


# ---------------------------------------------------------------------------
class ActivationTime2DTracker(Tracker):
    """act_t = where(act_t < 0 and u > threshold, t, act_t) on every grid node."""
    _native = True

    def __init__(self):
        super().__init__()
        self.act_t = np.ndarray
        self.threshold = -40
        self.file_name = "act_time_2d"
        self._dev = None

    def initialize(self, model):
        self.model = model
        self.act_t = -np.ones_like(self.model.u)
        self._dev = None

    def _track(self):   # host statement (only used if called by hand)
        self.act_t = np.where((self.act_t < 0) & (self.model.u > self.threshold),
                              self.model.t, self.act_t)

    def _register(self, engine, model, max_samples):
        if self._dev is None or self._dev.device != engine.device:
            self._dev = torch.from_numpy(np.ascontiguousarray(self.act_t, dtype=np.float64)).to(
                engine.device)
        self._id = engine.L.fwb_sim_add_tracker_act(
            engine.sim, ctypes.c_void_p(self._dev.data_ptr()), float(self.threshold),
            float(self.start_time), float(self.end_time), int(self.step))
        if self._id < 0:
            check(self._id, "fwb_sim_add_tracker_act")

    def _collect(self, engine):
        self.act_t = self._dev.cpu().numpy()

    @property
    def output(self):
        return self.act_t


class ActivationTime3DTracker(ActivationTime2DTracker):
    pass


class ECG2DTracker(Tracker):
    """Per lead sum over myocytes of (W u - u) / (d * dr), d = squared index
    distance (2D: (x-i)^2 + (y-j)^2 + z^2)."""
    _native = True

    def __init__(self, measure_coords=None):
        super().__init__()
        self.measure_coords = measure_coords
        self.ecg = []
        self.file_name = "ecg.npy"
        self.u_tr = None

    def initialize(self, model):
        self.model = model
        self.measure_coords = np.atleast_2d(self.measure_coords)
        if self.measure_coords.shape[1] != 3:
            raise ValueError("measure_coords must have 3 components (x, y, z) per lead")
        self.ecg = []
        self._taken = 0

    def calc_ecg(self):
        """The ECG signal of the model's current potential (reference: ecg_2d_tracker.py:61-79,
        ecg_3d_tracker.py:51-69), evaluated on the device by the stand-alone entry fwb_ecg:
        ``self.u_tr`` becomes W u on the myocytes, the return value is one number per lead."""
        model = self.model
        eng = getattr(model, "_engine", None)
        if eng is None or eng.weights is None:
            raise RuntimeError("calc_ecg() needs an initialised model (device weights)")
        dev = eng.device
        u = torch.from_numpy(np.ascontiguousarray(model.u, dtype=np.float64)).to(dev)
        u_tr = torch.zeros_like(u)
        coords = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(self.measure_coords),
                                                       dtype=np.float64)).to(dev)
        if coords.shape[1] != 3:
            raise ValueError("measure_coords must have 3 components (x, y, z) per lead")
        out = torch.zeros(coords.shape[0], dtype=torch.float64, device=dev)
        vp = ctypes.c_void_p
        check(eng.L.fwb_ecg(eng.dim, eng.stencil, shape_arr(eng.shape), vp(eng.chunk_bits.data_ptr()),
                            vp(eng.chunk_base.data_ptr()), eng.ld, vp(eng.worklist.data_ptr()),
                            eng.n_work, vp(u.data_ptr()), vp(u_tr.data_ptr()),
                            vp(eng.weights.data_ptr()), vp(coords.data_ptr()), int(coords.shape[0]),
                            float(model.dr), vp(out.data_ptr()),
                            vp(torch.cuda.current_stream().cuda_stream)), "fwb_ecg")
        self.u_tr = u_tr.cpu().numpy()
        return out.cpu().numpy()

    def _track(self):
        """By hand (the run loop samples inside the fused step kernel instead)."""
        self.ecg.append(self.calc_ecg())

    def _register(self, engine, model, max_samples):
        self._coords_dev = engine.keep(torch.from_numpy(np.ascontiguousarray(
            self.measure_coords, dtype=np.float64)).to(engine.device))
        self._n_leads = len(self.measure_coords)
        self._out = engine.keep(torch.zeros((max(1, max_samples), self._n_leads),
                                            dtype=torch.float64, device=engine.device))
        self._id = engine.L.fwb_sim_add_tracker_ecg(
            engine.sim, ctypes.c_void_p(self._coords_dev.data_ptr()), self._n_leads,
            float(model.dr), float(self.start_time), float(self.end_time), int(self.step),
            ctypes.c_void_p(self._out.data_ptr()), int(self._out.shape[0]))
        if self._id < 0:
            check(self._id, "fwb_sim_add_tracker_ecg")

    def _collect(self, engine):
        n = int(engine.L.fwb_sim_tracker_samples(engine.sim, self._id))
        rows = self._out[:n].cpu().numpy()
        self.ecg.extend(list(rows))

    @property
    def output(self):
        return np.array(self.ecg)

    def write(self):
        Path(self.path).mkdir(parents=True, exist_ok=True)
        np.save(Path(self.path).joinpath(self.file_name).with_suffix(".npy"), self.output)


class ECG3DTracker(ECG2DTracker):
    def write(self):
        Path(self.path).mkdir(parents=True, exist_ok=True)
        np.save(Path(self.path, self.file_name), self.output)


# ---- point samplers ------------------------------------------------------------
class _PointTracker(Tracker):
    _native = True

    def _vars(self):
        raise NotImplementedError

    def _append(self, rows):
        raise NotImplementedError

    def _register(self, engine, model, max_samples):
        cells = np.atleast_2d(self.cell_ind)
        flat = np.ravel_multi_index(tuple(cells.T), engine.shape)
        cidx = engine.compact_index(flat)
        names = ["u"] + [v for v in model.state_vars if v != "u"]
        items, fill = [], []
        for var in self._vars():
            if var not in names:
                raise ValueError(f"Variable '{var}' not found in model.")
            vid = names.index(var)
            host = model.__dict__.get(var)
            for f, c in zip(flat, cidx):
                items.append((vid, int(f), int(c)))
                fill.append(float(np.asarray(host).flat[int(f)]) if isinstance(host, np.ndarray)
                            else 0.0)
        self._n_cells = len(flat)
        self._items = engine.keep(torch.tensor(items, dtype=torch.int64, device=engine.device))
        self._fill = engine.keep(torch.tensor(fill, dtype=torch.float64, device=engine.device))
        self._out = engine.keep(torch.zeros((max(1, max_samples), len(items)),
                                            dtype=torch.float64, device=engine.device))
        self._id = engine.L.fwb_sim_add_tracker_point(
            engine.sim, ctypes.c_void_p(self._items.data_ptr()),
            ctypes.c_void_p(self._fill.data_ptr()), len(items), float(self.start_time),
            float(self.end_time), int(self.step), ctypes.c_void_p(self._out.data_ptr()),
            int(self._out.shape[0]))
        if self._id < 0:
            check(self._id, "fwb_sim_add_tracker_point")

    def _collect(self, engine):
        n = int(engine.L.fwb_sim_tracker_samples(engine.sim, self._id))
        self._append(self._out[:n].cpu().numpy())


class ActionPotential2DTracker(_PointTracker):
    def __init__(self):
        super().__init__()
        self.act_pot = []
        self.cell_ind = [1, 1]
        self.file_name = "act_pot"

    def initialize(self, model):
        self.model = model

    def _vars(self):
        return ["u"]

    def _append(self, rows):
        self.act_pot.extend(list(rows))

    def _track(self):
        cell = tuple(np.atleast_2d(self.cell_ind).T)
        self.act_pot.append(self.model.u[cell])

    @property
    def output(self):
        return np.squeeze(self.act_pot)


class ActionPotential3DTracker(ActionPotential2DTracker):
    pass


class MultiVariable2DTracker(_PointTracker):
    def __init__(self):
        super().__init__()
        self.var_list = []
        self.cell_ind = [1, 1]
        self.dir_name = "multi_vars"
        self.vars = {}

    def initialize(self, model):
        self.vars = {}
        self.model = model
        for var_ in self.var_list:
            if var_ not in self.model.__dict__:
                raise ValueError(f"Variable '{var_}' not found in model.")
            self.vars[var_] = []

    def _vars(self):
        return list(self.var_list)

    def _append(self, rows):
        n = self._n_cells
        for k, var_ in enumerate(self.var_list):
            self.vars[var_].extend(list(rows[:, k * n:(k + 1) * n]))

    def _track(self):
        cell = tuple(np.atleast_2d(self.cell_ind).T)
        for var_ in self.var_list:
            self.vars[var_].append(self.model.__dict__[var_][cell])

    @property
    def output(self):
        return {v: np.squeeze(self.vars[v]) for v in self.var_list}

    def write(self):
        out_dir = Path(self.path, self.dir_name)
        out_dir.mkdir(parents=True, exist_ok=True)
        for var_ in self.var_list:
            np.save(out_dir / f"{var_}.npy", self.output[var_])


class MultiVariable3DTracker(MultiVariable2DTracker):
    pass


class Variable2DTracker(MultiVariable2DTracker):
    @property
    def var_name(self):
        return self.var_list[0]

    @var_name.setter
    def var_name(self, value):
        self.var_list = [value]

    @property
    def output(self):
        return self.vars[self.var_name]

    def write(self):
        out_dir = Path(self.path, self.dir_name)
        out_dir.mkdir(parents=True, exist_ok=True)
        np.save(Path(out_dir, self.var_name).with_suffix(".npy"), self.output)


class Variable3DTracker(Variable2DTracker):
    pass


# ---- multi-activation trackers --------------------------------------
class LocalActivationTime2DTracker(Tracker):
    """Every threshold up-crossing of every node, in layers: a new layer is opened when a
    crossing node already holds a time in the current one (reference
    cpuwave2D/tracker/local_activation_time_2d_tracker.py:6-104).  Evaluated on the device
    (fwb_lat_cross / fwb_lat_write); ``output`` has shape ``(n_layers, *shape)``."""
    _device_hook = True

    def __init__(self):
        super().__init__()
        self.act_t = []
        self.threshold = -40
        self.file_name = "local_act_time_2d"
        self._activated = np.ndarray
        self._dev = None

    def initialize(self, model):
        self.model = model
        self.act_t = [-np.ones_like(self.model.u)]
        self._activated = np.full(self.model.u.shape, 0, dtype=bool)
        self._dev = None

    # device state: created from the host attributes on first use, so that values the user
    # assigned after initialize() are honoured
    def _ensure_dev(self, engine):
        if self._dev is not None and self._dev["device"] == engine.device:
            return self._dev
        dev = engine.device
        n = engine.n_nodes
        self._dev = dict(
            device=dev,
            activated=torch.from_numpy(np.ascontiguousarray(self._activated, dtype=np.uint8)
                                       .reshape(-1)).to(dev),
            cross=torch.zeros(n, dtype=torch.uint8, device=dev),
            flag=torch.zeros(1, dtype=torch.int32, device=dev),
            layers=[torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).to(dev)
                    for a in self.act_t],
            dirty=False)
        return self._dev

    def _track_device(self, engine, u, t):
        d = self._ensure_dev(engine)
        L, st = engine.L, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        n = engine.n_nodes
        check(L.fwb_lat_cross(ctypes.c_void_p(u.data_ptr()), n, float(self.threshold),
                              ctypes.c_void_p(d["activated"].data_ptr()),
                              ctypes.c_void_p(d["cross"].data_ptr()),
                              ctypes.c_void_p(d["layers"][-1].data_ptr()),
                              ctypes.c_void_p(d["flag"].data_ptr()), st), "fwb_lat_cross")
        if int(d["flag"].item()):          # 4-byte read-back: does this sample open a layer?
            d["layers"].append(torch.full((n,), -1.0, dtype=torch.float64, device=engine.device))
            d["flag"].zero_()
        check(L.fwb_lat_write(ctypes.c_void_p(d["cross"].data_ptr()), n, float(t),
                              ctypes.c_void_p(d["layers"][-1].data_ptr()), st), "fwb_lat_write")
        d["dirty"] = True

    def _track(self):
        eng = self.model._engine
        self._track_device(eng, eng.ubuf[eng.current()], self.model.t)

    def cross_threshold(self):
        """Public in the reference (local_activation_time_2d_tracker.py:71-90): the boolean
        array of the nodes of ``model.u`` that cross the threshold upwards now; arms / disarms
        the per-node ``activated`` state.  Called by hand it stages the host ``model.u`` on the
        device and runs fwb_lat_cross -- the same kernel the sampled steps use."""
        eng = self.model._engine
        d = self._ensure_dev(eng)
        u = torch.from_numpy(np.ascontiguousarray(self.model.u, dtype=np.float64)).to(eng.device)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(eng.L.fwb_lat_cross(ctypes.c_void_p(u.data_ptr()), eng.n_nodes,
                                  float(self.threshold),
                                  ctypes.c_void_p(d["activated"].data_ptr()),
                                  ctypes.c_void_p(d["cross"].data_ptr()), None, None, st),
              "fwb_lat_cross")
        d["dirty"] = True
        return d["cross"].cpu().numpy().reshape(self.model.u.shape).astype(bool)

    def _collect(self, engine=None):
        d = self._dev
        if d is None or not d["dirty"]:
            return
        shape = self.model.u.shape
        self.act_t = [a.cpu().numpy().reshape(shape) for a in d["layers"]]
        self._activated = d["activated"].cpu().numpy().reshape(shape).astype(bool)
        d["dirty"] = False

    @property
    def output(self):
        self._collect()
        return np.array(self.act_t)


class LocalActivationTime3DTracker(LocalActivationTime2DTracker):
    pass


class Period2DTracker(LocalActivationTime2DTracker):
    """Activation periods at detector cells (reference period_2d_tracker.py:9-86): the
    crossing detector runs over the whole grid on the device, only the detector cells' bits
    come back; ``output`` is a pandas DataFrame of the differences of successive activation
    times per cell, ``act_t`` the layers ``(n_layers, n_cells)``."""

    def __init__(self):
        super().__init__()
        self.cell_ind = []
        self.file_name = "period"

    def initialize(self, model):
        super().initialize(model)
        self.act_t = [-np.ones(len(np.atleast_2d(self.cell_ind)))]

    def _ensure_dev(self, engine):
        if self._dev is not None and self._dev["device"] == engine.device:
            return self._dev
        dev = engine.device
        cells = np.atleast_2d(self.cell_ind)
        flat = np.ravel_multi_index(tuple(cells.T), engine.shape)
        self._dev = dict(
            device=dev,
            activated=torch.from_numpy(np.ascontiguousarray(self._activated, dtype=np.uint8)
                                       .reshape(-1)).to(dev),
            cross=torch.zeros(engine.n_nodes, dtype=torch.uint8, device=dev),
            idx=torch.from_numpy(flat.astype(np.int64)).to(dev),
            out=torch.zeros(len(flat), dtype=torch.uint8, device=dev),
            dirty=False)
        return self._dev

    def _track_device(self, engine, u, t):
        d = self._ensure_dev(engine)
        L, st = engine.L, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(L.fwb_lat_cross(ctypes.c_void_p(u.data_ptr()), engine.n_nodes, float(self.threshold),
                              ctypes.c_void_p(d["activated"].data_ptr()),
                              ctypes.c_void_p(d["cross"].data_ptr()), None, None, st),
              "fwb_lat_cross")
        k = int(d["idx"].numel())
        check(L.fwb_gather_u8(ctypes.c_void_p(d["cross"].data_ptr()),
                              ctypes.c_void_p(d["idx"].data_ptr()), k,
                              ctypes.c_void_p(d["out"].data_ptr()), st), "fwb_gather_u8")
        cross = d["out"].cpu().numpy().astype(bool)
        if np.any(self.act_t[-1][cross] > -1):
            self.act_t.append(-np.ones(k))
        self.act_t[-1] = np.where(cross, t, self.act_t[-1])
        d["dirty"] = True

    def _collect(self, engine=None):
        d = self._dev
        if d is None or not d["dirty"]:
            return
        self._activated = d["activated"].cpu().numpy().reshape(self.model.u.shape).astype(bool)
        d["dirty"] = False

    @property
    def output(self):
        import pandas as pd
        lats = pd.DataFrame(np.array(self.act_t).T)
        return lats.apply(lambda row: np.diff(row[row != -1]), axis=1)

    def write(self):
        self.output.to_csv(Path(self.path, self.file_name).with_suffix(".csv"))


class Period3DTracker(Period2DTracker):
    pass


class SpiralWaveCore2DTracker(Tracker):
    """Spiral-wave tips: cells in which the threshold isoline of the previous and of the
    current sample cross, located by the bilinear interpolation of reference
    cpuwave2D/tracker/spiral_wave_core_2d_tracker.py:9-248.  The scan runs on the device
    (fwb_tip_scan); only the tips come back.  ``output`` is a DataFrame x, y, time, step."""
    _device_hook = True
    _COLS = ["x", "y", "time", "step"]

    def __init__(self):
        super().__init__()
        self.threshold = 0.5
        self.file_name = "spiral_wave_core"
        self.sprial_wave_cores = []          # (sic) the reference's attribute name
        self._dev = None

    def initialize(self, model):
        self.model = model
        self.u_prev = self.model.u.copy()
        self.sprial_wave_cores = []
        self._dev = None

    def _ensure_dev(self, engine):
        if self._dev is None or self._dev["device"] != engine.device:
            dev = engine.device
            self._dev = dict(
                device=dev, capacity=4096,
                u_prev=torch.from_numpy(np.ascontiguousarray(self.u_prev, dtype=np.float64)).to(dev),
                out=torch.zeros((4096, 3), dtype=torch.float64, device=dev),
                count=torch.zeros(1, dtype=torch.int32, device=dev))
        return self._dev

    def _scan(self, engine, u):
        d = self._ensure_dev(engine)
        L, st = engine.L, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        from ._lib import shape_arr
        while True:
            d["count"].zero_()
            check(L.fwb_tip_scan(ctypes.c_void_p(d["u_prev"].data_ptr()),
                                 ctypes.c_void_p(u.data_ptr()), len(engine.shape),
                                 shape_arr(engine.shape), float(self.threshold),
                                 ctypes.c_void_p(d["out"].data_ptr()), d["capacity"],
                                 ctypes.c_void_p(d["count"].data_ptr()), st), "fwb_tip_scan")
            n = int(d["count"].item())
            if n <= d["capacity"]:
                break
            d["capacity"] = 2 * n
            d["out"] = torch.zeros((d["capacity"], 3), dtype=torch.float64, device=engine.device)
        rows = d["out"][:n].cpu().numpy()
        rows = rows[np.argsort(rows[:, 2], kind="stable")]      # the reference's scan order
        d["u_prev"].copy_(u)
        return rows

    def track_tip_line(self, u, u_new, threshold):
        """Public in the reference (spiral_wave_core_2d_tracker.py:60-80): the tips between two
        2D potential fields as a list of ``[x, y]`` in the reference's scan order.  Stages the
        two host arrays on the device and runs fwb_tip_scan, the kernel the sampled steps use."""
        from ._lib import lib, shape_arr
        from .engine import require_cuda
        require_cuda()
        dev = torch.device(f"cuda:{torch.cuda.current_device()}")
        a = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64)).to(dev)
        b = torch.from_numpy(np.ascontiguousarray(u_new, dtype=np.float64)).to(dev)
        if a.dim() != 2 or a.shape != b.shape:
            raise ValueError("track_tip_line takes two 2D arrays of the same shape")
        L, st = lib(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        capacity = 4096
        while True:
            out = torch.zeros((capacity, 3), dtype=torch.float64, device=dev)
            count = torch.zeros(1, dtype=torch.int32, device=dev)
            check(L.fwb_tip_scan(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), 2,
                                 shape_arr(a.shape), float(threshold),
                                 ctypes.c_void_p(out.data_ptr()), capacity,
                                 ctypes.c_void_p(count.data_ptr()), st), "fwb_tip_scan")
            n = int(count.item())
            if n <= capacity:
                break
            capacity = 2 * n
        rows = out[:n].cpu().numpy()
        rows = rows[np.argsort(rows[:, 2], kind="stable")]
        return [[float(x), float(y)] for x, y in rows[:, :2]]

    def _frames(self, rows, t, step):
        import pandas as pd
        tips = pd.DataFrame(rows[:, :2], columns=["x", "y"])
        tips["time"] = t
        tips["step"] = step
        return [tips]

    def _track_device(self, engine, u, t):
        rows = self._scan(engine, u)
        self.sprial_wave_cores.extend(self._frames(rows, t, self.model.step))

    def _track(self):
        eng = self.model._engine
        self._track_device(eng, eng.ubuf[eng.current()], self.model.t)

    def write(self):
        self.output.to_csv(Path(self.path, self.file_name).with_suffix(".csv"))

    @property
    def output(self):
        import pandas as pd
        valid = [df for df in self.sprial_wave_cores if not df.empty]
        if not valid:
            return pd.DataFrame(columns=self._COLS)
        return pd.concat(valid, ignore_index=True)


class SpiralWaveCore3DTracker(SpiralWaveCore2DTracker):
    """The 2D scan on every slice of the last axis (reference
    cpuwave3D/tracker/spiral_wave_core_3d_tracker.py:7-48); columns x, y, z, time, step."""
    _COLS = ["x", "y", "z", "time", "step"]

    def _frames(self, rows, t, step):
        import pandas as pd
        n_i, n_j, n_k = self.model.u.shape
        z = (rows[:, 2] // (n_i * n_j)).astype(np.int64)
        frames = []
        for k in np.unique(z):                 # one frame per slice, like the reference
            sel = rows[z == k]
            tips = pd.DataFrame(sel[:, :2], columns=["x", "y"])
            tips["z"] = int(k)
            tips["time"] = t
            tips["step"] = step
            frames.append(tips)
        return frames


# ---- frame dumps ----------------------------------------------------
class Animation2DTracker(Tracker):
    """Saves ``model.<variable_name>`` as ``<path>/<dir_name>/<n>.npy`` at every sample
    (reference cpuwave2D/tracker/animation_2d_tracker.py:8-77).  The frame is snapshotted
    on the device and streamed out by hooks.FrameStreamer -- a copy stream and a writer
    thread -- so the step loop does not wait for the disk; all frames are on disk when
    ``run()`` returns.  ``write()`` (the mp4 builder of finitewave.tools) is not part of
    this backend."""
    _device_hook = True

    def __init__(self):
        super().__init__()
        self.dir_name = "animation"
        self.variable_name = "u"
        self.frame_type = "float64"
        self._frame_counter = 0
        self.overwrite = True
        self.file_name = "animation"
        self._streamer = None

    def initialize(self, model):
        self.model = model
        self._frame_counter = 0
        out = Path(self.path, self.dir_name)
        if not out.is_dir():
            out.mkdir(parents=True)
        if self.overwrite:
            for f in out.glob("*.npy"):
                f.unlink()
        self._streamer = None

    def _select(self):
        return None

    def _snapshot(self, engine, u):
        name = self.variable_name
        if name == "u":
            return u.clone()
        names = [v for v in self.model.state_vars if v != "u"]
        if name not in names:
            raise ValueError(f"Variable '{name}' not found in model.")
        slot = names.index(name)
        d = torch.empty(engine.shape, dtype=torch.float64, device=engine.device)
        from .engine import _ptr, _stream
        check(engine.L.fwb_scatter_compact(_ptr(engine.state[slot]), _ptr(d),
                                           float(self.model._init_value(name)), engine.n_nodes,
                                           _ptr(engine.chunk_bits), _ptr(engine.chunk_base),
                                           _stream()), "fwb_scatter_compact")
        return d

    def _track_device(self, engine, u, t):
        if self._streamer is None:
            from .hooks import FrameStreamer
            self._streamer = FrameStreamer(engine.shape)
        path = Path(self.path, self.dir_name, str(self._frame_counter)).with_suffix(".npy")
        self._streamer.submit(self._snapshot(engine, u), path, self.frame_type, self._select())
        self._frame_counter += 1

    def _track(self):
        """Called by hand (or from a host hook), outside the device loop: ``model.<variable>``
        is then a current host array and the reference's statement applies as it is
        (animation_2d_tracker.py:64-77, animation_slice_3d_tracker.py:45-56) -- a file write,
        no computation.  Works on any object with that attribute, as in the reference's tests."""
        frame = np.asarray(self.model.__dict__[self.variable_name])
        sel = self._select()
        if sel is not None:
            frame = frame[sel]
        np.save(Path(self.path, self.dir_name, str(self._frame_counter)).with_suffix(".npy"),
                frame.astype(self.frame_type))
        self._frame_counter += 1

    def _finish(self):
        """Called by run() before it returns: every submitted frame is on disk."""
        if self._streamer is not None:
            self._streamer.wait()

    def write(self, *args, **kwargs):
        raise NotImplementedError(
            "building the animation file is finitewave.tools' job (matplotlib / ffmpeg); the "
            f"frames are in {Path(self.path, self.dir_name)}")


class Animation3DTracker(Animation2DTracker):
    pass


class AnimationSlice3DTracker(Animation2DTracker):
    """One plane of a 3D variable per frame (reference
    cpuwave3D/tracker/animation_slice_3d_tracker.py:7-64): exactly one of ``slice_x``,
    ``slice_y``, ``slice_z`` must be set."""

    def __init__(self):
        super().__init__()
        self.slice_x = None
        self.slice_y = None
        self.slice_z = None

    def initialize(self, model):
        if np.count_nonzero([self.slice_x, self.slice_y, self.slice_z]) != 1:
            raise ValueError("Exactly one slice must be specified.")
        super().initialize(model)

    def _select(self):
        if self.slice_x is not None:
            return (self.slice_x, slice(None), slice(None))
        if self.slice_y is not None:
            return (slice(None), self.slice_y, slice(None))
        return (slice(None), slice(None), self.slice_z)

    def select_frame(self, array):
        """The 2D slice of a 3D array (animation_slice_3d_tracker.py:58-80)."""
        return array[self._select()]


class PeriodAnimation2DTracker(LocalActivationTime2DTracker):
    """Frame dump of the map of every node's latest threshold up-crossing time (reference
    cpuwave2D/tracker/period_animation_2d_tracker.py:7-72: ``period_map = where(cross, t,
    period_map)``, one ``.npy`` per sample).  Crossing detection and map update run on the
    device (fwb_lat_cross / fwb_lat_write); the frames are streamed out like the animation
    trackers' (hooks.FrameStreamer).  ``write()`` (the mp4 builder) is not part of this
    backend."""

    def __init__(self):
        super().__init__()
        self.dir_name = "period"
        self.file_name = "period"
        self.overwrite = False
        self._frame_counter = 0
        self.period_map = np.ndarray
        self._streamer = None

    def initialize(self, model):
        self.model = model
        self._frame_counter = 0
        self.period_map = -np.ones_like(self.model.u)
        self._activated = np.full(self.model.u.shape, 0, dtype=bool)
        out = Path(self.path, self.dir_name)
        if not out.is_dir():
            out.mkdir(parents=True)
        if self.overwrite:
            for f in out.glob("*.npy"):
                f.unlink()
        self._dev = None
        self._streamer = None

    def _ensure_dev(self, engine):
        if self._dev is None or self._dev["device"] != engine.device:
            dev = engine.device
            self._dev = dict(
                device=dev,
                activated=torch.from_numpy(np.ascontiguousarray(self._activated, dtype=np.uint8)
                                           .reshape(-1)).to(dev),
                cross=torch.zeros(engine.n_nodes, dtype=torch.uint8, device=dev),
                map=torch.from_numpy(np.ascontiguousarray(self.period_map, dtype=np.float64)).to(dev),
                dirty=False)
        return self._dev

    def _track_device(self, engine, u, t):
        d = self._ensure_dev(engine)
        L, st = engine.L, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        n = engine.n_nodes
        check(L.fwb_lat_cross(ctypes.c_void_p(u.data_ptr()), n, float(self.threshold),
                              ctypes.c_void_p(d["activated"].data_ptr()),
                              ctypes.c_void_p(d["cross"].data_ptr()), None, None, st), "fwb_lat_cross")
        check(L.fwb_lat_write(ctypes.c_void_p(d["cross"].data_ptr()), n, float(t),
                              ctypes.c_void_p(d["map"].data_ptr()), st), "fwb_lat_write")
        if self._streamer is None:
            from .hooks import FrameStreamer
            self._streamer = FrameStreamer(engine.shape)
        path = Path(self.path, self.dir_name, str(self._frame_counter)).with_suffix(".npy")
        self._streamer.submit(d["map"].clone(), path, "float64")
        self._frame_counter += 1
        d["dirty"] = True

    def _collect(self, engine=None):
        d = self._dev
        if d is None or not d["dirty"]:
            return
        shape = self.model.u.shape
        self.period_map = d["map"].cpu().numpy().reshape(shape)
        self._activated = d["activated"].cpu().numpy().reshape(shape).astype(bool)
        d["dirty"] = False

    def _finish(self):
        if self._streamer is not None:
            self._streamer.wait()
        self._collect()

    @property
    def output(self):
        self._collect()
        return self.period_map

    def write(self, *args, **kwargs):
        raise NotImplementedError(
            "building the animation file is finitewave.tools' job (matplotlib / ffmpeg); the "
            f"frames are in {Path(self.path, self.dir_name)}")


class PeriodAnimation3DTracker(PeriodAnimation2DTracker):
    pass
