"""
multi.py -- one tissue on several GPUs of ONE process, behind ``CardiacModel.run()``.

The reference is a single-process package; a user script that
builds a tissue, a model, stimuli and trackers and calls ``model.run()`` scales to the GPUs
of a box without edits:

    FWB_DEVICES=all python my_script.py          # or  model.devices = [0, 1, 2, 3]

``MultiEngine`` presents the interface ``CardiacModel`` uses from ``Engine`` and spreads it
over one ``Engine`` per device: the tissue is cut into slabs along axis 0 (slab.py), every
device stores its owned slices plus one ghost slice per neighbour, the step kernels'
slab-boundary blocks store ``u_new`` straight into the neighbour's ghost slice through a peer
mapping (NVLink) and raise its flag -- the in-kernel halo exchange of the multi-process
runs (include/finitewave_b200.h, "Slab decomposition"), wired with raw pointers instead of
IPC handles.  The steps are enqueued by one C call, round-robin over the slabs in short
chunks (fwb_multi_run): a slab's boundary blocks never wait for a neighbour launch that is
stuck behind a full launch queue, and no host thread is in the middle of an allocation or a
free -- both synchronise the device -- while a kernel spins on a neighbour's flag.

Host arrays stay whole: ``model.u`` and the state arrays are the reference's dense numpy
arrays over the full tissue; a slab uploads / downloads its slices of them (axis-0 slices
of a C-ordered array are contiguous, so these are plain async copies from / to the pinned
arrays).  Coordinate-box stimuli, ActivationTime and ECG trackers run on the devices; every
other hook (Commands, savers, user-defined trackers / stimuli, point samplers) runs on the
host between device segments exactly as with one GPU.
"""
import copy
import ctypes

import numpy as np
import torch

from . import _lib, slab
from ._lib import check
from .engine import Engine


def requested_devices(model):
    """Device indices asked for by ``model.devices`` or FWB_DEVICES, or None."""
    import os
    want = getattr(model, "devices", None)
    if want is None:
        want = os.environ.get("FWB_DEVICES")
    if want is None or want == "":
        return None
    if isinstance(want, str):
        if want.strip().lower() == "all":
            want = list(range(torch.cuda.device_count()))
        else:
            want = [int(x) for x in want.split(",") if x.strip() != ""]
    want = [int(d) for d in want]
    return want if len(want) >= 2 else None


def supported(model, shape, devices):
    """(ok, reason): can this model / tissue run slab-decomposed in one process?"""
    if min(devices) < 0 or max(devices) >= torch.cuda.device_count():
        return False, "a requested device is not visible"
    if shape[-1] % 32 != 0:
        return False, "the contiguous axis must be a multiple of 32 nodes"
    if int(np.prod(shape[1:])) % 32 != 0 or shape[0] < 4 * len(devices):
        return False, "needs at least 4 slices of axis 0 per device"
    for st in (model.stim_sequence.sequence if model.stim_sequence else []):
        if getattr(st, "_native", False) and not hasattr(st, "_box"):
            return False, f"{type(st).__name__} is not a coordinate-box stimulus"
    for tr in (model.tracker_sequence.sequence if model.tracker_sequence else []):
        if getattr(tr, "_device_hook", False):
            return False, f"{type(tr).__name__} runs its own device kernels (single GPU only)"
    return True, ""


class MultiEngine:
    """Device side of one CardiacModel on several GPUs (see the module docstring)."""

    multi = True

    def __init__(self, shape, devices):
        self.L = _lib.lib()
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        self.devices = [int(d) for d in devices]
        n = len(self.devices)
        self.owned = slab.partition(self.shape[0], n)
        self.stored = [slab.stored_range(o, self.shape[0]) for o in self.owned]
        self.engines = []
        for d, (lo, hi, halo) in zip(self.devices, self.stored):
            with torch.cuda.device(d):
                self.engines.append(Engine((hi - lo, *self.shape[1:]), device=f"cuda:{d}"))
        # one stream per slab: two slabs may share a device (tests on a one-GPU box), and a
        # slab's boundary blocks wait for flags that its neighbour's kernels raise
        self.streams = []
        for d in self.devices:
            with torch.cuda.device(d):
                self.streams.append(torch.cuda.Stream(device=d))
        self.device = self.engines[0].device
        self._created = False
        self._keep = []
        self.n_state = 0
        self._native = None
        for a in self.devices:
            for b in self.devices:
                if a != b and abs(self.devices.index(a) - self.devices.index(b)) == 1:
                    check(self.L.fwb_enable_peer_access(a, b), "fwb_enable_peer_access")

    # ---- helpers ----------------------------------------------------------------
    def _each(self, fn):
        """fn(rank, engine) on every device, under that device's context; list of results."""
        out = []
        for r, (d, e) in enumerate(zip(self.devices, self.engines)):
            with torch.cuda.device(d), torch.cuda.stream(self.streams[r]):
                out.append(fn(r, e))
        return out

    def _rows(self, r):
        """(owned range in global slices, owned range in the slab's stored slices)"""
        a, b = self.owned[r]
        lo = self.stored[r][0]
        return (a, b), (a - lo, b - lo)

    def _slice(self, r, arr):
        lo, hi, _ = self.stored[r]
        return arr[lo:hi]

    # ---- tissue / weights ---------------------------------------------------------
    def set_tissue(self, mesh, special_boundaries=None, halo=None):
        mesh = np.ascontiguousarray(mesh)
        sb = None if special_boundaries is None else np.ascontiguousarray(special_boundaries)
        self._each(lambda r, e: e.set_tissue(self._slice(r, mesh),
                                             None if sb is None else self._slice(r, sb),
                                             halo=self.stored[r][2]))

    @property
    def n_myo(self):
        return sum(e.n_myo for e in self.engines)

    @property
    def ld(self):
        return tuple(e.ld for e in self.engines)

    @property
    def K(self):
        return self.engines[0].K

    @property
    def stencil(self):
        return self.engines[0].stencil

    @property
    def weights(self):
        return self.engines[0].weights

    def compute_weights(self, stencil, conductivity, fibers, D_al, D_ac, D_model, dt, dr):
        def one(r, e):
            c = conductivity
            if not isinstance(c, torch.Tensor) and np.ndim(c) != 0:
                c = self._slice(r, np.asarray(c) * np.ones(self.shape))
            f = None if fibers is None else self._slice(r, np.asarray(fibers))
            e.compute_weights(stencil, c, f, D_al, D_ac, D_model, dt, dr)
        self._each(one)

    def set_weights_dense(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        self._each(lambda r, e: e.set_weights_dense(self._slice(r, w)))

    def weights_dense(self):
        parts = self._each(lambda r, e: e.weights_dense())
        out = np.empty((*self.shape, self.K), dtype=np.float64)
        for r, p in enumerate(parts):
            (a, b), (la, lb) = self._rows(r)
            out[a:b] = p[la:lb]
        return out

    # ---- buffers ------------------------------------------------------------------
    def needs_allocation(self, n_state):
        return any(e.needs_allocation(n_state) for e in self.engines)

    def allocate(self, n_state, staging=True, peer=True):
        self.n_state = n_state
        self._each(lambda r, e: e.allocate(n_state, staging=True, peer=True))

    def upload_dense(self, which, host):
        self._each(lambda r, e: e.upload_dense(which, self._slice(r, host)))

    def download_dense(self, which, host_out):
        def one(r, e):
            (a, b), loc = self._rows(r)
            e.download_dense(which, host_out[a:b], rows=loc)
        self._each(one)

    def upload_state(self, slot, host, fill=None):
        def one(r, e):
            (a, b), loc = self._rows(r)
            e.upload_state(slot, host[a:b], fill=fill, rows=loc)
        self._each(one)

    def update_copy_idle(self):
        self._each(lambda r, e: e.update_copy_idle())

    def off_fill(self):
        per = [e.off_fill() for e in self.engines]
        return [sum(p[i] for p in per if i < len(p)) for i in range(max(len(p) for p in per))] \
            if per and any(per) else []

    def download_state(self, slot, host_out, fill, keep=False):
        def one(r, e):
            (a, b), loc = self._rows(r)
            e.download_state(slot, host_out[a:b], fill, keep=keep, rows=loc)
        self._each(one)

    # ---- simulation objects -----------------------------------------------------------
    @property
    def sim(self):
        return self._created

    def create_sim(self, model_id, params, dt, use_tma=True):
        self.destroy_sim()

        def one(r, e):
            e.create_sim(model_id, params, dt, use_tma=use_tma)
            lo = self.stored[r][0]
            if lo:
                check(self.L.fwb_sim_set_slow_offset(e.sim, lo), "fwb_sim_set_slow_offset")
        self._each(one)
        # halos: raw pointers (one address space), peer access enabled in __init__
        vp = ctypes.c_void_p

        def side(e):
            if e is None:
                return vp(0), vp(0), 0, vp(0)
            return vp(e.peer_mem[0].ptr), vp(e.peer_mem[1].ptr), e.shape[0], vp(e.peer_flags.ptr)

        def wire(r, e):
            halo = self.stored[r][2]
            lo = side(self.engines[r - 1] if halo[0] else None)
            hi = side(self.engines[r + 1] if halo[1] else None)
            n_lo, n_hi = e.halo_blocks
            check(self.L.fwb_sim_set_halo(e.sim, vp(e.peer_flags.ptr), lo[0], lo[1], lo[2], lo[3],
                                          n_lo, hi[0], hi[1], hi[2], hi[3], n_hi),
                  "fwb_sim_set_halo")
        self._each(wire)
        self._created = True
        self._keep = []
        self._native = None

    def destroy_sim(self):
        if self._created:
            self.synchronize()
        self._each(lambda r, e: e.destroy_sim())
        self._created = False

    def set_params(self, p, dt):
        arr = (ctypes.c_double * len(p))(*p)
        self._each(lambda r, e: check(self.L.fwb_sim_set_params(e.sim, arr, len(p), float(dt)),
                                      "fwb_sim_set_params"))

    def current(self):
        return self.engines[0].current()

    def set_time(self, t, step):
        self._each(lambda r, e: e.set_time(t, step))

    def get_time(self):
        return self.engines[0].get_time()

    def run(self, n_steps):
        """Every slab advances n_steps: one C call enqueues them round-robin in short chunks
        on the slabs' streams (fwb_multi_run)."""
        n = len(self.engines)
        arr = (ctypes.c_void_p * n)(*[e.sim.value for e in self.engines])
        check(self.L.fwb_multi_run(arr, n, int(n_steps), 8), "fwb_multi_run")

    def launch_count(self):
        return sum(e.launch_count() for e in self.engines) if self._created else 0

    def device_steps(self):
        return self.engines[0].device_steps() if self._created else 0

    def keep(self, t):
        self._keep.append(t)
        return t

    def synchronize(self):
        self._each(lambda r, e: e.synchronize())

    # ---- native stimuli / trackers ---------------------------------------------------
    def register_native(self, model, live):
        """Coordinate-box stimuli and ActivationTime / ECG trackers on every slab."""
        from .tracker import ActivationTime2DTracker, ECG2DTracker
        self._each(lambda r, e: (check(self.L.fwb_sim_clear_stims(e.sim)),
                                 check(self.L.fwb_sim_clear_trackers(e.sim))))
        for e in self.engines:
            e._keep = []
        native = dict(stims=[], act=[], ecg=[])
        if live["native_stims"]:
            for st in live["stims"]:
                ids = []
                for r, e in enumerate(self.engines):
                    lo, hi, _ = self.stored[r]
                    loc = copy.copy(st)
                    loc.x1, loc.x2 = slab.global_box_to_local(st.x1, st.x2, self.shape[0], lo, hi - lo)
                    with torch.cuda.device(self.devices[r]), torch.cuda.stream(self.streams[r]):
                        sid = loc._register(e, model)
                        check(self.L.fwb_sim_set_stim_passed(e.sim, sid, int(bool(st.passed))))
                    ids.append(sid)
                native["stims"].append((st, ids))
        remaining = max(0, live["iters"] - live["done"])
        for tr in live["native_tr"]:
            cap = remaining // max(1, int(tr.step)) + 2
            if isinstance(tr, ActivationTime2DTracker):
                devs = []
                for r, e in enumerate(self.engines):
                    with torch.cuda.device(self.devices[r]), torch.cuda.stream(self.streams[r]):
                        t = torch.from_numpy(np.ascontiguousarray(
                            self._slice(r, np.asarray(tr.act_t)), dtype=np.float64)).to(e.device)
                        rc = self.L.fwb_sim_add_tracker_act(
                            e.sim, ctypes.c_void_p(t.data_ptr()), float(tr.threshold),
                            float(tr.start_time), float(tr.end_time), int(tr.step))
                        if rc < 0:
                            check(rc, "fwb_sim_add_tracker_act")
                        devs.append(e.keep(t))
                native["act"].append((tr, devs))
            elif isinstance(tr, ECG2DTracker):
                coords = np.ascontiguousarray(tr.measure_coords, dtype=np.float64)
                outs = []
                for r, e in enumerate(self.engines):
                    with torch.cuda.device(self.devices[r]), torch.cuda.stream(self.streams[r]):
                        c = e.keep(torch.from_numpy(coords).to(e.device))
                        o = e.keep(torch.zeros((max(1, cap), len(coords)), dtype=torch.float64,
                                               device=e.device))
                        rc = self.L.fwb_sim_add_tracker_ecg(
                            e.sim, ctypes.c_void_p(c.data_ptr()), len(coords), float(model.dr),
                            float(tr.start_time), float(tr.end_time), int(tr.step),
                            ctypes.c_void_p(o.data_ptr()), int(o.shape[0]))
                        if rc < 0:
                            check(rc, "fwb_sim_add_tracker_ecg")
                        outs.append((rc, o))
                native["ecg"].append((tr, outs))
            else:
                raise RuntimeError(f"{type(tr).__name__} cannot run slab-decomposed")
        self._native = native

    def collect_native(self, model, live):
        self.synchronize()
        nat = self._native
        if nat is None:
            return
        for st, ids in nat["stims"]:
            rc = self.L.fwb_sim_stim_passed(self.engines[0].sim, ids[0])
            if rc < 0:
                raise _lib.FwbError("fwb_sim_stim_passed failed on slab 0")
            st.passed = bool(rc)
        for tr, devs in nat["act"]:
            act = np.array(tr.act_t, dtype=np.float64, copy=True)
            for r, t in enumerate(devs):
                (a, b), (la, lb) = self._rows(r)
                act[a:b] = t[la:lb].cpu().numpy()
            tr.act_t = act
        for tr, outs in nat["ecg"]:
            n = int(self.L.fwb_sim_tracker_samples(self.engines[0].sim, outs[0][0]))
            if n > 0:
                # lead sums are sums over nodes: add the slabs' partial sums in slab order
                total = outs[0][1][:n].cpu().numpy().copy()
                for tid, o in outs[1:]:
                    total += o[:n].cpu().numpy()
                tr.ecg.extend(list(total))
        self._native = None
