"""
DeviceSimulation -- a simulation whose tissue and state live only on the GPU.

``CardiacModel.run()`` mirrors the reference's host-array contract (numpy arrays in
``model.__dict__``), which caps the tissue size at what the host can hold.  The
large configurations (512^3 ... 1024^3, SURVEY.md section 8d "generated on device per
slab") use this class instead: mesh / fibres / conductivity are torch CUDA tensors
(or numpy arrays that fit), state arrays are filled with the model's ``init_*``
constants directly on the device, stimuli and trackers are the same objects as in
the host API, and ``run(n_steps)`` advances the fused step kernel with the
reference's loop order (cardiac_model.py:164-189).  Nothing ever round-trips through
host memory except what the caller asks for (``u_host()``, ``state_host(name)``,
tracker outputs).

``model`` is an un-run CardiacModel instance: it supplies the model id, parameter
values, ``init_*`` constants, ``dt``, ``dr``, ``D_model`` and optionally ``stencil``.
"""
import copy
import ctypes

import numpy as np
import torch

from . import _lib
from .engine import Engine


class _Shim:
    """What the stimulus / tracker ``_register`` hooks need from a model."""

    def __init__(self, sim):
        self._sim = sim
        self.dt = sim.dt
        self.dr = sim.dr
        self.state_vars = ["u"] + list(sim.state_names)
        self.cardiac_tissue = sim

    @property
    def mesh(self):
        return self._sim.mesh_host()


class DeviceSimulation:
    def __init__(self, model, mesh, fibers=None, conductivity=1.0, device=None, stencil=None,
                 halo=(False, False), slow_offset=0, global_slices=None):
        """halo / slow_offset / global_slices describe a slab of a larger tissue cut along
        axis 0 (see slab.py): ``mesh`` then includes one ghost slice per neighbour."""
        self.model = model
        shape = tuple(int(s) for s in mesh.shape)
        if len(shape) != model._DIM:
            raise ValueError(f"{type(model).__name__} needs a {model._DIM}-dimensional mesh")
        self.shape = shape
        self.dt, self.dr = float(model.dt), float(model.dr)
        self.state_names = list(model._STATE)
        self.engine = eng = Engine(shape, device)
        self._mesh = mesh
        self.halo = (bool(halo[0]), bool(halo[1]))
        self.slow_offset = int(slow_offset)
        self.global_slices = int(global_slices) if global_slices is not None else shape[0]
        eng.set_tissue(mesh, halo=self.halo)
        if stencil is None:
            stencil = model.stencil
        # the stencil's own kind (a SymmetricStencil2D is an AsymmetricStencil2D subclass
        # with different weights); without a stencil object: by the presence of fibres
        if stencil is None:
            kind = _lib.STENCIL_ISO if fibers is None else _lib.STENCIL_ANISO
        else:
            kind = getattr(stencil, "_KIND", None)
            if kind is None:
                raise ValueError("DeviceSimulation takes the built-in Stencil classes only")
        aniso = kind != _lib.STENCIL_ISO
        if aniso and fibers is None:
            raise ValueError("Fibers must be provided for anisotropic diffusion.")
        D_al = getattr(stencil, "D_al", 1) if stencil is not None else 1
        D_ac = getattr(stencil, "D_ac", 1 / 9) if stencil is not None else 1 / 9
        eng.compute_weights(kind, conductivity,
                            fibers if aniso else None, D_al, D_ac, model.D_model, self.dt, self.dr)
        eng.allocate(len(self.state_names), staging=False, peer=any(self.halo))
        eng.ubuf[0].fill_(float(model.init_u))
        eng.ubuf[1].fill_(float(model.init_u) if model._INIT_U_NEW else 0.0)
        for slot, name in enumerate(self.state_names):
            eng.fill_state(slot, model._init_value(name))
        eng.create_sim(_lib.MODEL_IDS[model._MODEL], model._param_vector(), self.dt)
        if self.slow_offset:
            _lib.check(eng.L.fwb_sim_set_slow_offset(eng.sim, self.slow_offset))
        eng.update_copy_idle()
        self.t, self.step = 0.0, 0
        self.stims, self.trackers = [], []
        self._shim = _Shim(self)

    # ---- tissue views ----------------------------------------------------
    @property
    def n_myo(self):
        return self.engine.n_myo

    @property
    def mesh(self):
        return self.mesh_host()

    def mesh_host(self):
        m = self._mesh
        return m.cpu().numpy() if isinstance(m, torch.Tensor) else np.asarray(m)

    # ---- stimuli / trackers (same classes as the host API) -----------------
    def add_stim(self, stim):
        """Stimulus boxes are given in GLOBAL tissue indices; a slab registers the part
        that falls into its stored slices (ghost slices included, so both neighbours
        apply the identical edit to their copy of a shared slice)."""
        stim.initialize(self._shim)
        if any(self.halo) or self.slow_offset:
            if not hasattr(stim, "_box"):
                raise NotImplementedError("slab runs take coordinate-box stimuli only")
            stim = copy.copy(stim)
            lo, hi, _ = slice(stim.x1, stim.x2).indices(self.global_slices)
            lo, hi = lo - self.slow_offset, max(lo, hi) - self.slow_offset
            stim.x1 = min(max(lo, 0), self.shape[0])
            stim.x2 = min(max(hi, stim.x1), self.shape[0])
        sid = stim._register(self.engine, self._shim)
        self.stims.append((sid, stim))
        return stim

    def add_tracker(self, tracker, max_samples):
        if not getattr(tracker, "_native", False):
            raise ValueError("DeviceSimulation takes native trackers only")
        tracker.model = self._shim
        if hasattr(tracker, "measure_coords"):
            tracker.initialize(self._shim)
        elif hasattr(tracker, "act_t"):
            tracker._dev = torch.full(self.shape, -1.0, dtype=torch.float64,
                                      device=self.engine.device)
        tracker._register(self.engine, self._shim, int(max_samples))
        self.trackers.append(tracker)
        return tracker

    # ---- time loop -----------------------------------------------------------
    def run(self, n_steps, halo_sync=True):
        eng = self.engine
        eng.set_time(self.t, self.step)
        eng.run(int(n_steps))
        if any(self.halo) and halo_sync:
            _lib.check(eng.L.fwb_sim_halo_sync(eng.sim), "fwb_sim_halo_sync")
        self.t, self.step = eng.get_time()

    # ---- slab wiring -----------------------------------------------------------
    def halo_export(self):
        """What a neighbour needs to reach this slab's buffers from another process."""
        eng = self.engine
        return dict(u0=eng.peer_mem[0].ipc_handle(), u1=eng.peer_mem[1].ipc_handle(),
                    flags=eng.peer_flags.ipc_handle(), slices=self.shape[0])

    def halo_pointers(self):
        """Same, as raw device pointers (neighbour slab in the same process)."""
        eng = self.engine
        return dict(u0=eng.peer_mem[0].ptr, u1=eng.peer_mem[1].ptr, flags=eng.peer_flags.ptr,
                    slices=self.shape[0])

    def halo_connect(self, lo=None, hi=None):
        """lo / hi: dict(u0, u1, flags, slices) with device pointers valid in this process."""
        eng = self.engine
        vp = ctypes.c_void_p

        def side(d):
            if d is None:
                return vp(0), vp(0), 0, vp(0)
            return vp(d["u0"]), vp(d["u1"]), int(d["slices"]), vp(d["flags"])
        l, h = side(lo), side(hi)
        n_lo, n_hi = eng.halo_blocks
        _lib.check(eng.L.fwb_sim_set_halo(eng.sim, vp(eng.peer_flags.ptr), l[0], l[1], l[2], l[3],
                                          n_lo, h[0], h[1], h[2], h[3], n_hi), "fwb_sim_set_halo")

    def synchronize(self):
        self.engine.synchronize()

    def launch_count(self):
        return self.engine.launch_count()

    # ---- results ---------------------------------------------------------------
    def u_device(self):
        return self.engine.ubuf[self.engine.current()]

    def u_host(self):
        return self.u_device().cpu().numpy()

    def state_compact(self, name):
        """Device view [n_myo] of one state variable in the device's compact (tile)
        order; use state_host() for the dense array."""
        return self.engine.state[self.state_names.index(name), :self.engine.n_myo]

    def state_host(self, name):
        eng = self.engine
        out = torch.empty(self.shape, dtype=torch.float64, device=eng.device)
        _lib.check(eng.L.fwb_scatter_compact(
            ctypes.c_void_p(eng.state[self.state_names.index(name)].data_ptr()),
            ctypes.c_void_p(out.data_ptr()), float(self.model._init_value(name)),
            eng.n_nodes, ctypes.c_void_p(eng.chunk_bits.data_ptr()),
            ctypes.c_void_p(eng.chunk_base.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "fwb_scatter_compact")
        return out.cpu().numpy()

    # ---- host round trip of the whole state (slab-local, stored range) -----------
    def host_buffers(self):
        """Pinned host arrays for download_host / upload_host: u, u_new and every state
        variable over this simulation's stored range."""
        from .engine import pinned_empty
        return {name: pinned_empty(self.shape) for name in ["u", "u_new"] + self.state_names}

    def download_host(self, bufs):
        eng = self.engine
        cur = eng.current()
        bufs["u"].copy_(eng.ubuf[cur], non_blocking=True)
        bufs["u_new"].copy_(eng.ubuf[cur ^ 1], non_blocking=True)
        for slot, name in enumerate(self.state_names):
            eng.download_state(slot, bufs[name], self.model._init_value(name))
        eng.synchronize()

    def upload_host(self, bufs):
        eng = self.engine
        cur = eng.current()
        eng.ubuf[cur].copy_(bufs["u"], non_blocking=True)
        eng.ubuf[cur ^ 1].copy_(bufs["u_new"], non_blocking=True)
        for slot, name in enumerate(self.state_names):
            eng.upload_state(slot, bufs[name])
        eng.update_copy_idle()
        eng.synchronize()

    def collect(self):
        self.engine.synchronize()
        for tr in self.trackers:
            tr._collect(self.engine)
