"""
CardiacModel and the six on-path models, 2D and 3D.

Public surface = the reference's (finitewave/core/model/cardiac_model.py:9-239 and
finitewave/cpuwave{2D,3D}/model/{aliev_panfilov,barkley,mitchell_schaeffer,fenton_karma,
luo_rudy91,tp06}_{2,3}d.py): attribute-style configuration (``dt, dr, t_max, prog_bar,
npfloat, D_model``, every model parameter and ``init_*``), ``cardiac_tissue /
stim_sequence / tracker_sequence / command_sequence / state_loader / state_saver /
stencil``, ``u``, ``u_new``, each state variable, ``weights``, ``t``, ``step``,
``initialize()``, ``compute_weights()``, ``run(initialize=True, num_of_theads=None)``,
``check_termination()``, ``select_stencil()``, ``clone()``.

What differs is where the time loop runs: ``run()`` keeps the reference's step
order but executes stimulus -> diffusion -> ionic -> native
trackers -> ``t += dt`` on the GPU, as one fused kernel per step, for as many steps
as may pass before the next *host hook* (a Command whose time has come, a
StateSaver that fires, a user-defined Tracker whose gate passes, a user-defined
Stim that is due, or termination).  Around every host hook the ``state_vars``
arrays in ``model.__dict__`` are synchronised with the device, so hooks see and
may mutate plain numpy arrays exactly as with the reference.
"""
import copy
import ctypes

import numpy as np

from . import _lib
from .engine import Engine, pinned_empty
from .stencil import (AsymmetricStencil2D, AsymmetricStencil3D, DeviceWeights,
                      IsotropicStencil2D, IsotropicStencil3D)


class CardiacModel:
    # subclasses fill these
    _MODEL = None          # key of _lib.MODEL_IDS
    _PARAMS = ()           # parameter attribute names in the ionic kernel's argument order
    _STATE = ()            # state attribute names in device slot order
    _INIT_U_NEW = False    # LR91/TP06 also initialise u_new (luo_rudy91_2d.py:117, tp06_2d.py:182)
    _DIM = None

    def __init__(self):
        self.meta = {}
        self.cardiac_tissue = None
        self.stim_sequence = None
        self.tracker_sequence = None
        self.command_sequence = None
        self.state_loader = None
        self.state_saver = None
        self.stencil = None
        self.diffusion_kernel = None
        self.ionic_kernel = None
        self.u = np.ndarray
        self.u_new = np.ndarray
        self.weights = np.ndarray
        self.dt = 0.
        self.dr = 0.
        self.t_max = 0.
        self.t = 0
        self.step = 0
        self.D_model = 1.
        self.prog_bar = True
        self.npfloat = np.float64
        self.state_vars = []
        self._engine = None
        self._engine_key = None
        self.gpu_launches = 0       # kernels launched by the last run()
        self.gpu_steps = 0          # time steps the device kernels advanced in the last run()
        self.devices = None         # e.g. [0, 1, 2, 3] or "all": slabs on several GPUs (multi.py);
                                    # None: FWB_DEVICES, else the current device

    # ------------------------------------------------------------------ engine
    def _engine_for(self, tissue):
        """The device side: one Engine, or -- when ``model.devices`` / FWB_DEVICES names
        several GPUs and the tissue can be cut into slabs -- a MultiEngine (multi.py)."""
        from . import multi
        shape = tuple(tissue.mesh.shape)
        devices = multi.requested_devices(self)
        want_multi = False
        if devices:
            ok, why = multi.supported(self, shape, devices)
            if ok:
                want_multi = True
            elif getattr(self, "_multi_note", None) != why:
                self._multi_note = why
                import warnings
                warnings.warn(f"finitewave_b200: running on one GPU ({why})")
        cur = self._engine
        if (cur is None or cur.shape != shape or bool(getattr(cur, "multi", False)) != want_multi
                or (want_multi and cur.devices != devices)):
            if cur is not None:
                cur.destroy_sim()
            self._engine = multi.MultiEngine(shape, devices) if want_multi else Engine(shape)
        return self._engine

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_engine",):
                new.__dict__[k] = None
            elif isinstance(v, DeviceWeights):
                new.__dict__[k] = np.asarray(v).copy()
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def clone(self):
        return copy.deepcopy(self)

    # ------------------------------------------------------------------ setup
    def _init_value(self, name):
        """init_<var> of a state attribute (Courtemanche's `j_` array is filled from
        `init_j`, courtemanche_2d.py:126)."""
        return getattr(self, "init_" + name.rstrip("_"))

    def _alloc_host(self, shape, value):
        a = pinned_empty(shape).numpy()
        a[...] = value
        return a

    def initialize(self):
        """CardiacModel.initialize (cardiac_model.py:91-117) + the subclass part that
        fills u and the state arrays with their init_* constants."""
        if np.dtype(self.npfloat) != np.float64:
            raise NotImplementedError("finitewave_b200 computes in float64 only")
        mesh = self.cardiac_tissue.mesh
        shape = mesh.shape
        self.u = self._alloc_host(shape, 0.0)
        self.u_new = self._alloc_host(shape, 0.0)
        self.step = 0
        self.t = 0

        self.compute_weights()
        self.diffusion_kernel = self.stencil.select_diffusion_kernel()

        for seq in (self.stim_sequence, self.tracker_sequence, self.command_sequence,
                    self.state_loader, self.state_saver):
            if seq:
                seq.initialize(self)

        self.u[...] = self.init_u
        if self._INIT_U_NEW:
            self.u_new[...] = self.init_u
        for name in self._STATE:
            self.__dict__[name] = self._alloc_host(shape, self._init_value(name))
        self._uploaded = False

    def compute_weights(self):
        """cardiac_model.py:119-128; may be called again by a Command after the mesh
        changed (Tutorials/SpiralWaves2D.ipynb UpdateMesh)."""
        tissue = self.cardiac_tissue
        tissue.compute_myo_indexes()
        if self.stencil is None:
            self.stencil = self.select_stencil(tissue)
        eng = self._engine_for(tissue)
        if eng.sim:
            # every index structure of the device simulation is about to be replaced
            # (new mask -> new chunks, work list, compact order): keep what the native
            # trackers sampled so far, drop the simulation object, rebuild on upload
            if getattr(self, "_live", None):
                self._collect_native()
            eng.destroy_sim()
        eng.set_tissue(tissue.mesh, tissue.special_boundaries)
        w = self.stencil.compute_weights(self, tissue)
        if not isinstance(w, DeviceWeights):
            eng.set_weights_dense(np.asarray(w))
            w = DeviceWeights(eng)
        self.weights = w

    def _rebuild_device_side(self):
        """run(initialize=False) on a model without a device side -- a clone()
        (``copy.deepcopy`` keeps the host arrays and drops the engine): tissue index structures
        again, and the weights the model carries (a clone holds them as the reference's
        ``(*shape, K)`` ndarray) packed back onto the device, so the run continues exactly
        where the copied arrays stand."""
        tissue = self.cardiac_tissue
        if tissue is None or not isinstance(self.__dict__.get("u"), np.ndarray):
            raise RuntimeError("run(initialize=False) needs an initialised model")
        w = self.weights
        if isinstance(w, np.ndarray) and w.shape[:-1] == tuple(tissue.mesh.shape):
            tissue.compute_myo_indexes()
            if self.stencil is None:
                self.stencil = self.select_stencil(tissue)
            eng = self._engine_for(tissue)
            eng.set_tissue(tissue.mesh, tissue.special_boundaries)
            eng.set_weights_dense(w)
            self.weights = DeviceWeights(eng)
        else:
            self.compute_weights()

    def select_stencil(self, cardiac_tissue):
        iso, aniso = ((IsotropicStencil2D, AsymmetricStencil2D) if self._DIM == 2
                      else (IsotropicStencil3D, AsymmetricStencil3D))
        return iso() if cardiac_tissue.fibers is None else aniso()

    # ------------------------------------------------------------------ kernels
    def run_diffusion_kernel(self):
        """u_new = W u on the host-visible arrays (reference cardiac_model.py:205-211)."""
        self.diffusion_kernel(self.u_new, self.u, self.weights, self.cardiac_tissue.myo_indexes)

    def run_ionic_kernel(self):
        raise NotImplementedError(
            "finitewave_b200 fuses diffusion and the ionic update into one device kernel; "
            "use run() (or Engine.run) instead of calling the two halves separately.")

    # ------------------------------------------------------------------ device sync
    def _param_vector(self):
        return [float(getattr(self, p)) for p in self._PARAMS]

    def _upload(self):
        """Push u, u_new and every state array from model.__dict__ to the device
        (creating / re-creating the device simulation when needed)."""
        eng = self._engine
        if eng.needs_allocation(len(self._STATE)):
            eng.destroy_sim()
            eng.allocate(len(self._STATE))
        if not eng.sim:
            eng.create_sim(_lib.MODEL_IDS[self._MODEL], self._param_vector(), self.dt)
        else:
            eng.set_params(self._param_vector(), self.dt)
        cur = eng.current()
        eng.upload_dense(cur, self._host_array("u"))
        eng.upload_dense(cur ^ 1, self._host_array("u_new"))
        for slot, name in enumerate(self._STATE):
            eng.upload_state(slot, self._host_array(name), fill=self._init_value(name))
        eng.synchronize()
        # state rows only exist for updated nodes; where the host arrays hold something
        # else than init_* on the other nodes, downloads must preserve it
        self._keep_offnodes = [c != 0 for c in eng.off_fill()]
        eng.update_copy_idle()

    def _partition(self):
        """Who handles what, from the sequences as they are NOW (the reference re-reads them
        every step, so a Command may add, remove, re-arm or replace stimuli and trackers):
        built-in stimuli / trackers run on the device, the rest are host hooks."""
        stims = list(self.stim_sequence.sequence) if self.stim_sequence else []
        native_stims = bool(self.stim_sequence) and self.stim_sequence.all_native()
        trackers = list(self.tracker_sequence.sequence) if self.tracker_sequence else []
        native_tr = [tr for tr in trackers if getattr(tr, "_native", False)]
        if getattr(self._engine, "multi", False):
            # slab-decomposed: the field trackers run on the devices, point samplers (a few
            # scalars of one slab) are sampled on the host like user-defined trackers
            from .tracker import ActivationTime2DTracker, ECG2DTracker
            native_tr = [tr for tr in native_tr
                         if isinstance(tr, (ActivationTime2DTracker, ECG2DTracker))]
        dev_tr = [tr for tr in trackers if getattr(tr, "_device_hook", False)]
        host_tr = [tr for tr in trackers if tr not in native_tr
                   and not getattr(tr, "_device_hook", False)]
        commands = self.command_sequence.sequence if self.command_sequence else []
        self._live.update(stims=stims, native_stims=native_stims, native_tr=native_tr)
        return stims, native_stims, native_tr, dev_tr, host_tr, commands

    def _register_native(self):
        """(Re-)register the built-in stimuli and trackers with the device runner."""
        eng, live = self._engine, self._live
        live["registered"] = True
        if getattr(eng, "multi", False):
            eng.register_native(self, live)
            return
        _lib.check(eng.L.fwb_sim_clear_stims(eng.sim))
        _lib.check(eng.L.fwb_sim_clear_trackers(eng.sim))
        eng._keep = []
        if live["native_stims"]:
            for st in live["stims"]:
                sid = st._register(eng, self)
                _lib.check(eng.L.fwb_sim_set_stim_passed(eng.sim, sid, int(bool(st.passed))))
        remaining = max(0, live["iters"] - live["done"])
        for tr in live["native_tr"]:
            tr._register(eng, self, remaining // max(1, int(tr.step)) + 2)

    def _collect_native(self):
        """Samples and stimulus flags from the device into the host objects -- once per
        registration (a Command that calls compute_weights() right after the loop has
        collected must not append the same samples twice)."""
        eng, live = self._engine, self._live
        if not live.get("registered") or not eng.sim:
            return
        live["registered"] = False
        if getattr(eng, "multi", False):
            eng.collect_native(self, live)
            return
        eng.synchronize()
        for tr in live["native_tr"]:
            tr._collect(eng)
        if live["native_stims"]:
            for i, st in enumerate(live["stims"]):
                rc = eng.L.fwb_sim_stim_passed(eng.sim, i)
                if rc < 0:
                    raise _lib.FwbError(f"fwb_sim_stim_passed({i}) failed (rc={rc}): stimulus "
                                        "sequence and device registration are out of step")
                st.passed = bool(rc)
                if hasattr(st, "_collect"):
                    st._collect(eng, i)      # StimVoltageListMatrix3D: its own step counter

    def _host_array(self, name):
        a = self.__dict__[name]
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != self._engine.shape:
            raise ValueError(f"model.{name} has shape {a.shape}, expected {self._engine.shape}")
        return a

    def _download(self, swapped_view=False):
        """Copy u, u_new and every state array from the device into the numpy arrays
        of model.__dict__.  swapped_view: expose the pre-swap assignment of the two
        potential buffers (what a tracker sees mid-step)."""
        eng = self._engine
        cur = eng.current()
        if swapped_view:
            cur ^= 1
        for name, which in (("u", cur), ("u_new", cur ^ 1)):
            host = self.__dict__[name]
            if not (isinstance(host, np.ndarray) and host.dtype == np.float64
                    and host.flags.c_contiguous and host.shape == eng.shape):
                host = np.empty(eng.shape, dtype=np.float64)
                self.__dict__[name] = host
            eng.download_dense(which, host)
        keep_flags = getattr(self, "_keep_offnodes", [])
        for slot, name in enumerate(self._STATE):
            host = self.__dict__[name]
            keep = slot < len(keep_flags) and keep_flags[slot]
            if not (isinstance(host, np.ndarray) and host.dtype == np.float64
                    and host.flags.c_contiguous and host.shape == eng.shape):
                host = np.array(host, dtype=np.float64, order="C") if keep else \
                    np.empty(eng.shape, dtype=np.float64)
                self.__dict__[name] = host
            eng.download_state(slot, host, self._init_value(name), keep=keep)
        eng.synchronize()

    # ------------------------------------------------------------------ the loop
    def check_termination(self):
        max_iters = int(np.ceil(self.t_max / self.dt))
        return (self.t > self.t_max) or (self.step > max_iters)

    def run(self, initialize=True, num_of_theads=None):
        """Runs the simulation (reference: cardiac_model.py:130-189).  `num_of_theads`
        is accepted for signature compatibility; the work runs on the GPU."""
        if initialize:
            self.initialize()
        if self.t_max < self.t:
            raise ValueError("t_max must be greater than current t.")
        if self.state_loader:
            self.state_loader.load()

        iters = int(np.ceil((self.t_max - self.t) / self.dt))
        if self._engine is None:
            self._rebuild_device_side()
        eng = self._engine

        self._live = dict(iters=iters, done=0)
        stims, native_stims, native_tr, dev_tr, host_tr, commands = self._partition()
        self._upload()
        self._register_native()
        launches0 = eng.launch_count()
        dsteps0 = eng.device_steps()
        eng.set_time(self.t, self.step)

        bar = None
        if self.prog_bar:
            from tqdm import tqdm
            bar = tqdm(total=iters, desc=f"Running {self.__class__.__name__}")
        chunk_cap = max(1, iters // 100) if self.prog_bar else max(1, iters)

        done = 0
        host_view_valid = True      # model.__dict__ arrays == device content
        finished = False
        while done < iters and not finished:
            # ---- plan: how many steps until the next host hook --------------------
            t_sim, step_sim, n = self.t, self.step, 0
            hook_stim = hook_tracker = hook_dev = False
            while done + n < iters and n < chunk_cap:
                due_s = (not native_stims) and any(t_sim >= s.t and not s.passed for s in stims)
                due_t = any(tr.gate(t_sim, step_sim) for tr in host_tr)
                due_d = any(tr.gate(t_sim, step_sim) for tr in dev_tr)
                if (due_s or due_t or due_d) and n > 0:
                    break                  # run the n clean steps first
                t_sim += self.dt
                step_sim += 1
                n += 1
                if due_s or due_t or due_d:
                    hook_stim, hook_tracker, hook_dev = due_s, due_t, due_d
                    break                  # this single step carries the hook
                if self._post_hook_due(t_sim, step_sim, commands):
                    break
            post_hook = n > 0 and self._post_hook_due(t_sim, step_sim, commands)

            # ---- pre-step host hook: user-defined stimuli --------------------------
            if hook_stim:
                if not host_view_valid:
                    self._download()
                self.stim_sequence.stimulate_next()
                eng.upload_dense(eng.current(), self._host_array("u"))
                eng.update_copy_idle()
                host_view_valid = False

            # ---- the device steps ---------------------------------------------------
            eng.set_time(self.t, self.step)
            eng.run(n)
            host_view_valid = False

            # ---- mid-step device hook: trackers with their own device kernels; they see
            # the step's pre-update potential (the buffer that was `u` before the swap)
            if hook_dev:
                for tr in dev_tr:
                    if tr.gate(self.t, self.step):
                        tr._track_device(eng, eng.ubuf[eng.current() ^ 1], self.t)

            # ---- mid-step host hook: user-defined trackers (pre-increment view) ------
            if hook_tracker:
                self._download(swapped_view=True)
                for tr in host_tr:
                    tr.track()
                # restore the post-swap assignment of the two potential arrays
                d = self.__dict__
                d["u"], d["u_new"] = d["u_new"], d["u"]
                host_view_valid = True

            self.t, self.step = t_sim, step_sim
            done += n
            self._live["done"] = done
            if bar is not None:
                bar.update(n)

            # ---- post-step host hooks: commands, state savers, termination -----------
            if post_hook:
                if any((not c.passed) and self.t >= c.t for c in commands):
                    if not host_view_valid:
                        self._download()
                    # host objects first catch up with the device (samples taken so far,
                    # which stimuli have passed), then the command runs, then everything it
                    # may have touched -- arrays, parameters, the stimulus and tracker
                    # sequences themselves -- goes back to the device
                    if eng.sim:
                        self._collect_native()
                    self.command_sequence.execute_next()
                    stims, native_stims, native_tr, dev_tr, host_tr, commands = self._partition()
                    self._upload()
                    self._register_native()
                    host_view_valid = True
                if self.state_saver and self._saver_due(self.t):
                    if not self._save_async(host_view_valid):
                        if not host_view_valid:
                            self._download()
                            host_view_valid = True
                        self.state_saver.save()
                if self.check_termination():
                    if self.state_saver:
                        if not host_view_valid:
                            self._download()
                            host_view_valid = True
                        self.state_saver.save()
                    finished = True

        if bar is not None:
            bar.close()
        if not host_view_valid:
            self._download()
        self._collect_native()
        if getattr(self, "_ckpt_writer", None) is not None:
            self._ckpt_writer.wait()          # checkpoint files are complete on return
        for tr in dev_tr:
            if hasattr(tr, "_finish"):
                tr._finish()                  # streamed frames are complete on return
        self.gpu_launches = eng.launch_count() - launches0
        self.gpu_steps = eng.device_steps() - dsteps0
        self._live = None

    def _save_async(self, host_view_valid):
        """Mid-run checkpoint without stalling the step loop.  Applies to
        the built-in StateSaver class (a subclass may override how variables are written) when
        the device holds the only current copy and no state array carries user values on
        nodes the solver does not update.  Returns False to request the synchronous path."""
        from .hooks import AsyncCheckpointWriter, StateSaver, StateSaverCollection
        if host_view_valid or getattr(self, "async_checkpoints", True) is False:
            return False
        if getattr(self._engine, "multi", False):
            return False
        sv = self.state_saver
        savers = list(sv.savers) if isinstance(sv, StateSaverCollection) else [sv]
        due = [x for x in savers if x.due()]
        if not due or any(type(x) is not StateSaver for x in due):
            return False
        if any(getattr(self, "_keep_offnodes", [])) or self.t >= self.t_max:
            return False
        if list(self.state_vars) != ["u"] + list(self._STATE):
            return False
        eng = self._engine
        fills = [self._init_value(name) for name in self._STATE]
        arrays, done, keep = eng.snapshot_async(fills)
        if getattr(self, "_ckpt_writer", None) is None:
            self._ckpt_writer = AsyncCheckpointWriter()
        for x in due:
            self._ckpt_writer.submit(x.path, list(self.state_vars), arrays, done, keep)
            x.passed = True
        return True

    def _saver_due(self, t):
        savers = getattr(self.state_saver, "savers", None)
        savers = savers if savers is not None and len(savers) else [self.state_saver]
        for sv in savers:
            if sv.passed:
                continue
            if sv.time < 0 and t >= self.t_max:
                return True
            if sv.time >= 0 and t >= sv.time:
                return True
        return False

    def _post_hook_due(self, t, step, commands):
        if any((not c.passed) and t >= c.t for c in commands):
            return True
        if self.state_saver and self._saver_due(t):
            return True
        max_iters = int(np.ceil(self.t_max / self.dt))
        return (t > self.t_max) or (step > max_iters)


# ---------------------------------------------------------------------------
# The six models.  Defaults and attribute names are the reference's.
# ---------------------------------------------------------------------------
class AlievPanfilov2D(CardiacModel):
    """aliev_panfilov_2d.py:13-131"""
    _MODEL, _DIM = "aliev_panfilov", 2
    _PARAMS = ("a", "k", "eap", "mu_1", "mu_2")
    _STATE = ("v",)

    def __init__(self):
        super().__init__()
        self.v = np.ndarray
        self.D_model = 1.
        self.state_vars = ["u", "v"]
        self.npfloat = 'float64'
        self.a, self.k, self.eap, self.mu_1, self.mu_2 = 0.1, 8.0, 0.01, 0.2, 0.3
        self.init_u, self.init_v = 0.0, 0.0


class Barkley2D(CardiacModel):
    """barkley_2d.py:13-110"""
    _MODEL, _DIM = "barkley", 2
    _PARAMS = ("a", "b", "eap")
    _STATE = ("v",)

    def __init__(self):
        super().__init__()
        self.v = np.ndarray
        self.D_model = 1.
        self.state_vars = ["u", "v"]
        self.npfloat = 'float64'
        self.a, self.b, self.eap = 0.75, 0.02, 0.02
        self.init_u, self.init_v = 0.0, 0


class MitchellSchaeffer2D(CardiacModel):
    """mitchell_schaeffer_2d.py:13-100"""
    _MODEL, _DIM = "mitchell_schaeffer", 2
    _PARAMS = ("tau_close", "tau_open", "tau_in", "tau_out", "u_gate")
    _STATE = ("h",)

    def __init__(self):
        super().__init__()
        self.h = np.ndarray
        self.D_model = 1.
        self.state_vars = ["u", "h"]
        self.npfloat = 'float64'
        self.tau_close, self.tau_open, self.tau_out, self.tau_in = 150, 120, 6, 0.3
        self.u_gate = 0.13
        self.init_u, self.init_h = 0.0, 1.0


class FentonKarma2D(CardiacModel):
    """fenton_karma_2d.py:13-141 (MLR-I parameter set)"""
    _MODEL, _DIM = "fenton_karma", 2
    _PARAMS = ("tau_d", "tau_o", "tau_r", "tau_si", "tau_v_m", "tau_v_p", "tau_w_m",
               "tau_w_p", "k", "u_c", "uc_si")
    _STATE = ("v", "w")

    def __init__(self):
        super().__init__()
        self.v = np.ndarray
        self.w = np.ndarray
        self.D_model = 1.
        self.state_vars = ["u", "v", "w"]
        self.npfloat = 'float64'
        self.tau_r, self.tau_o, self.tau_d, self.tau_si = 130, 12.5, 0.172, 127
        self.tau_v_m, self.tau_v_p, self.tau_w_m, self.tau_w_p = 18.2, 10, 80, 1020
        self.k, self.u_c, self.uc_si = 10, 0.13, 0.85
        self.init_u, self.init_v, self.init_w = 0.0, 1.0, 1.0


class BuenoOrovio2D(CardiacModel):
    """bueno_orovio_2d.py:13-181 (minimal ventricular model)"""
    _MODEL, _DIM = "bueno_orovio", 2
    _PARAMS = ("u_o", "u_u", "theta_v", "theta_w", "theta_v_m", "theta_o", "tau_v1_m",
               "tau_v2_m", "tau_v_p", "tau_w1_m", "tau_w2_m", "k_w_m", "u_w_m", "tau_w_p",
               "tau_fi", "tau_o1", "tau_o2", "tau_so1", "tau_so2", "k_so", "u_so", "tau_s1",
               "tau_s2", "k_s", "u_s", "tau_si", "tau_w_inf", "w_inf_")
    _STATE = ("v", "w", "s")

    def __init__(self):
        super().__init__()
        self.v = np.ndarray
        self.w = np.ndarray
        self.s = np.ndarray
        self.D_model = 1.
        self.state_vars = ["u", "v", "w", "s"]
        self.npfloat = 'float64'
        self.u_o, self.u_u = 0.0, 1.55
        self.theta_v, self.theta_w, self.theta_v_m, self.theta_o = 0.3, 0.13, 0.006, 0.006
        self.tau_v1_m, self.tau_v2_m, self.tau_v_p = 60, 1150, 1.4506
        self.tau_w1_m, self.tau_w2_m, self.k_w_m, self.u_w_m, self.tau_w_p = 60, 15, 65, 0.03, 200
        self.tau_fi, self.tau_o1, self.tau_o2 = 0.11, 400, 6
        self.tau_so1, self.tau_so2, self.k_so, self.u_so = 30.0181, 0.9957, 2.0458, 0.65
        self.tau_s1, self.tau_s2, self.k_s, self.u_s = 2.7342, 16, 2.0994, 0.9087
        self.tau_si, self.tau_w_inf, self.w_inf_ = 1.8875, 0.07, 0.94
        self.init_u, self.init_v, self.init_w, self.init_s = 0.0, 1.0, 1.0, 0.0


class Courtemanche2D(CardiacModel):
    """courtemanche_2d.py:12-178 (human atrial model)"""
    _MODEL, _DIM = "courtemanche", 2
    _PARAMS = ("gna", "gnab", "gk1", "gkr", "gks", "gto", "gcal", "gcab", "gkur_coeff", "F", "T",
               "R", "Vc", "Vj", "Vup", "Vrel", "ibk", "cao", "nao", "ko", "caupmax", "kup",
               "kmnai", "kmko", "kmnancx", "kmcancx", "ksatncx", "kmcmdn", "kmtrpn", "kmcsqn",
               "trpnmax", "cmdnmax", "csqnmax", "inacamax", "inakmax", "ipcamax", "krel",
               "iupmax", "kq10")
    _STATE = ("nai", "ki", "cai", "caup", "carel", "m", "h", "j_", "d", "f", "oa", "oi", "ua",
              "ui", "xr", "xs", "fca", "irel", "vrel", "urel", "wrel")
    _INIT_U_NEW = True

    def __init__(self):
        super().__init__()
        self.D_model = 0.154
        self.state_vars = ["u"] + list(self._STATE)
        self.npfloat = 'float64'
        self.gna, self.gnab, self.gk1, self.gkr, self.gks = 7.8, 0.000674, 0.09, 0.0294, 0.129
        self.gto, self.gcal, self.gcab, self.gkur_coeff = 0.1652, 0.1238, 0.00113, 1
        self.F, self.T, self.R = 96485.0, 310.0, 8314.0
        self.Vc = 20100
        self.Vj = self.Vc * 0.68
        self.Vup = self.Vj * 0.06 * 0.92
        self.Vrel = self.Vj * 0.06 * 0.08
        self.ibk, self.cao, self.nao, self.ko = 0.0, 1.8, 140, 5.4
        self.caupmax, self.kup, self.kmnai, self.kmko = 15, 0.00092, 10, 1.5
        self.kmnancx, self.kmcancx, self.ksatncx = 87.5, 1.38, 0.1
        self.kmcmdn, self.kmtrpn, self.kmcsqn = 0.00238, 0.0005, 0.8
        self.trpnmax, self.cmdnmax, self.csqnmax = 0.07, 0.05, 10.0
        self.inacamax, self.inakmax, self.ipcamax = 1600, 0.6, 0.275
        self.krel, self.iupmax, self.kq10 = 30, 0.005, 3
        self.init_u = -84.5
        self.init_nai, self.init_ki, self.init_cai = 11.2, 139, 0.000102
        self.init_caup, self.init_carel = 1.6, 1.1
        self.init_m, self.init_h, self.init_j = 0.00291, 0.965, 0.978
        self.init_d, self.init_f = 0.000137, 0.999837
        self.init_oa, self.init_oi, self.init_ua, self.init_ui = 0.000592, 0.9992, 0.003519, 0.9987
        self.init_xs, self.init_xr, self.init_fca = 0.0187, 0.0000329, 0.775
        self.init_irel, self.init_vrel, self.init_urel, self.init_wrel = 0, 1, 0, 0.9


class LuoRudy912D(CardiacModel):
    """luo_rudy91_2d.py:12-156"""
    _MODEL, _DIM = "luo_rudy91", 2
    _PARAMS = ("gna", "gsi", "gk", "gk1", "gkp", "gb", "ko", "ki", "nai", "nao", "cao",
               "R", "T", "F", "PR_NaK")
    _STATE = ("m", "h", "j", "d", "f", "x", "cai")
    _INIT_U_NEW = True

    def __init__(self):
        super().__init__()
        self.D_model = 0.1
        for name in self._STATE:
            setattr(self, name, np.ndarray)
        self.state_vars = ["u", "m", "h", "j", "d", "f", "x", "cai"]
        self.npfloat = 'float64'
        self.gna, self.gsi, self.gk, self.gk1 = 23.0, 0.09, 0.282, 0.6047
        self.gkp, self.gb = 0.0183, 0.03921
        self.ko, self.ki, self.nai, self.nao, self.cao = 5.4, 145.0, 18.0, 140.0, 1.8
        self.R, self.T, self.F, self.PR_NaK = 8.314, 310.0, 96.5, 0.01833
        self.init_u = -84.5
        self.init_m, self.init_h, self.init_j = 0.0017, 0.9832, 0.995484
        self.init_d, self.init_f, self.init_x, self.init_cai = 0.000003, 1.0, 0.0057, 0.0002


class TP062D(CardiacModel):
    """tp06_2d.py:13-240 (the reference does not pre-declare the state attributes,
    so e.g. MultiVariable trackers on TP06 states fail at initialize there; here
    they are declared after the first initialize() as well)"""
    _MODEL, _DIM = "tp06", 2
    _PARAMS = ("ko", "cao", "nao", "Vc", "Vsr", "Vss", "Bufc", "Kbufc", "Bufsr", "Kbufsr",
               "Bufss", "Kbufss", "Vmaxup", "Kup", "Vrel", "k1_", "k2_", "k3", "k4", "EC",
               "maxsr", "minsr", "Vleak", "Vxfer", "R", "F", "T", "RTONF", "CAPACITANCE",
               "gkr", "pKNa", "gk1", "gna", "gbna", "KmK", "KmNa", "knak", "gcal", "gbca",
               "knaca", "KmNai", "KmCa", "ksat", "n_", "gpca", "KpCa", "gpk", "gto", "gks")
    _STATE = ("cai", "casr", "cass", "nai", "Ki", "m", "h", "j", "xr1", "xr2", "xs", "r",
              "s", "d", "f", "f2", "fcass", "rr", "oo")
    _INIT_U_NEW = True

    def __init__(self):
        super().__init__()
        self.D_model = 0.154
        self.state_vars = ["u", "cai", "casr", "cass", "nai", "Ki", "m", "h", "j", "xr1",
                           "xr2", "xs", "r", "s", "d", "f", "f2", "fcass", "rr", "oo"]
        self.npfloat = 'float64'
        self.ko, self.cao, self.nao = 5.4, 2.0, 140.0
        self.Vc, self.Vsr, self.Vss = 0.016404, 0.001094, 0.00005468
        self.Bufc, self.Kbufc, self.Bufsr, self.Kbufsr = 0.2, 0.001, 10., 0.3
        self.Bufss, self.Kbufss = 0.4, 0.00025
        self.Vmaxup, self.Kup, self.Vrel = 0.006375, 0.00025, 0.102
        self.k1_, self.k2_, self.k3, self.k4 = 0.15, 0.045, 0.060, 0.005
        self.EC, self.maxsr, self.minsr = 1.5, 2.5, 1.
        self.Vleak, self.Vxfer = 0.00036, 0.0038
        self.R, self.F, self.T, self.RTONF = 8314.472, 96485.3415, 310.0, 26.71376
        self.CAPACITANCE = 0.185
        self.gkr, self.gks, self.gk1, self.gto, self.gna = 0.153, 0.392, 5.405, 0.294, 14.838
        self.gbna, self.gcal, self.gbca, self.gpca = 0.00029, 0.00003980, 0.000592, 0.1238
        self.KpCa, self.gpk, self.pKNa = 0.0005, 0.0146, 0.03
        self.KmK, self.KmNa, self.knak, self.knaca = 1.0, 40.0, 2.724, 1000
        self.KmNai, self.KmCa, self.ksat, self.n_ = 87.5, 1.38, 0.1, 0.35
        self.init_u = -84.5
        self.init_cai, self.init_casr, self.init_cass = 0.00007, 1.3, 0.00007
        self.init_nai, self.init_Ki = 7.67, 138.3
        self.init_m, self.init_h, self.init_j = 0., 0.75, 0.75
        self.init_xr1, self.init_xr2, self.init_xs = 0., 1., 0.
        self.init_r, self.init_s = 0., 1.
        self.init_d, self.init_f, self.init_f2, self.init_fcass = 0., 1., 1., 1.
        self.init_rr, self.init_oo = 1., 0.


def _as_3d(cls2d, name):
    return type(name, (cls2d,), {"_DIM": 3, "__doc__": f"3D twin of {cls2d.__name__} "
                                 "(the reference's 3D classes only re-index the same point functions)"})


AlievPanfilov3D = _as_3d(AlievPanfilov2D, "AlievPanfilov3D")
Barkley3D = _as_3d(Barkley2D, "Barkley3D")
MitchellSchaeffer3D = _as_3d(MitchellSchaeffer2D, "MitchellSchaeffer3D")
FentonKarma3D = _as_3d(FentonKarma2D, "FentonKarma3D")
BuenoOrovio3D = _as_3d(BuenoOrovio2D, "BuenoOrovio3D")
Courtemanche3D = _as_3d(Courtemanche2D, "Courtemanche3D")
LuoRudy913D = _as_3d(LuoRudy912D, "LuoRudy913D")
TP063D = _as_3d(TP062D, "TP063D")
