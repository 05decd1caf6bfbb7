"""
Test double for finitewave_b200.engine.Engine (TEST INFRASTRUCTURE, CPU only).

`OracleEngine` offers the interface `CardiacModel.run()` talks to, but keeps every buffer
as a numpy array and advances the time steps with the CPU oracle (oracle/).  It exists so
that the `-m "not gpu"` suite can run the product's HOST logic -- the planning of device
segments around host hooks, the stimuli's host faces, the firing / gating rules, commands,
state savers and loaders, the float accumulation of `t` -- against the golden fixtures of the
live reference without a GPU.  Built-in stimuli and trackers are forced onto their host faces
(`host_faces_only`), because their native faces register descriptors with the CUDA library.

Nothing under finitewave_b200/ imports this file (tests/test_cabi_exports.py checks the
product never reaches the oracle); the GPU parity tests never use it.
"""
import numpy as np

from oracle import oracle

_KINDS = {0: "iso", 1: "aniso", 2: "sym"}


class _FakeLib:
    """The handful of C-ABI entry points model.py calls directly on `engine.L`."""

    def fwb_sim_set_params(self, sim, arr, n, dt):
        sim.params = np.array([arr[i] for i in range(n)], dtype=np.float64)
        sim.dt = float(dt)
        return 0

    def fwb_sim_clear_stims(self, sim):
        return 0

    def fwb_sim_clear_trackers(self, sim):
        return 0

    def fwb_last_error(self):
        return b""


class _Sim:
    def __init__(self, model, params, dt):
        self.model, self.params, self.dt = model, np.array(params, dtype=np.float64), float(dt)
        self.cur, self.t, self.step, self.launches = 0, 0, 0, 0


class OracleEngine:
    device = "cpu"

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        self.n_nodes = int(np.prod(self.shape))
        self.L = _FakeLib()
        self.sim = None
        self._keep = []
        self.state = None
        self.ubuf = [None, None]
        self.weights = None
        self.K = 0
        self.runs = []          # lengths of the device segments run() was asked for

    # ---- tissue / weights -------------------------------------------------
    def set_tissue(self, mesh, special_boundaries=None, halo=(False, False)):
        self.destroy_sim()
        self.mesh = np.array(mesh)
        on = self.mesh == 1
        if special_boundaries is not None:
            on &= np.asarray(special_boundaries) == 0
        self.indexes = np.flatnonzero(on).astype(np.int64)
        self.n_myo = len(self.indexes)
        self.ld = max(32, (self.n_myo + 31) // 32 * 32)

    def compute_weights(self, stencil, conductivity, fibers, D_al, D_ac, D_model, dt, dr):
        self.kind = _KINDS[stencil]
        self.weights = oracle.compute_weights(self.mesh, conductivity, fibers, self.kind,
                                              D_model, dt, dr, D_al=D_al, D_ac=D_ac)
        self.K = self.weights.shape[-1]
        self.off = oracle.flat_offsets(self.kind, self.shape)

    def set_weights_dense(self, w):
        self.weights = np.ascontiguousarray(w, dtype=np.float64)
        self.K = self.weights.shape[-1]
        self.kind = "iso" if self.K in (5, 7) else "aniso"
        self.off = oracle.flat_offsets(self.kind, self.shape)

    def weights_dense(self):
        return self.weights

    # ---- buffers -----------------------------------------------------------
    def allocate(self, n_state, staging=True, peer=False):
        self.n_state = n_state
        self.ubuf = [np.zeros(self.shape), np.zeros(self.shape)]
        self.dense_state = [np.zeros(self.shape) for _ in range(n_state)]
        self.state = np.zeros((max(n_state, 1), self.ld))     # only its shape is looked at
        self._off = [0] * n_state

    def update_copy_idle(self):
        pass

    def needs_allocation(self, n_state):
        return (self.state is None or self.ubuf[0] is None
                or self.state.shape != (max(n_state, 1), self.ld))

    def upload_dense(self, which, host):
        self.ubuf[which][...] = host

    def download_dense(self, which, host_out):
        host_out[...] = self.ubuf[which]

    def upload_state(self, slot, host, fill=None):
        self.dense_state[slot][...] = host
        if fill is not None:
            off = np.ones(self.n_nodes, dtype=bool)
            off[self.indexes] = False
            self._off[slot] = int(np.count_nonzero(np.asarray(host).reshape(-1)[off] != fill))

    def off_fill(self):
        return list(self._off)

    def download_state(self, slot, host_out, fill, keep=False):
        flat = host_out.reshape(-1)
        if not keep:
            flat[...] = fill
        flat[self.indexes] = self.dense_state[slot].reshape(-1)[self.indexes]

    def snapshot_async(self, fills):
        """Same contract as Engine.snapshot_async: (host arrays [u, state...], an event whose
        synchronize() says they are valid, keep-alive list)."""
        arrays = [self.ubuf[self.sim.cur].copy()]
        for slot in range(self.n_state):
            a = np.full(self.shape, float(fills[slot]))
            self.download_state(slot, a, fills[slot])
            arrays.append(a)

        class _Done:
            def synchronize(self):
                pass
        return arrays, _Done(), list(arrays)

    # ---- simulation object ---------------------------------------------------
    def create_sim(self, model_id, params, dt, use_tma=True):
        from finitewave_b200 import _lib
        names = {v: k for k, v in _lib.MODEL_IDS.items()}
        self.sim = _Sim(names[model_id], params, dt)
        self._keep = []

    def destroy_sim(self):
        self.sim = None

    def set_params(self, p, dt):
        self.sim.params = np.array([float(x) for x in p], dtype=np.float64)
        self.sim.dt = float(dt)

    def current(self):
        return self.sim.cur

    def set_time(self, t, step):
        self.sim.t, self.sim.step = t, int(step)

    def run(self, n_steps):
        s = self.sim
        self.runs.append(int(n_steps))
        for _ in range(int(n_steps)):
            u, u_new = self.ubuf[s.cur], self.ubuf[s.cur ^ 1]
            oracle.diffuse(u_new, u, self.weights, self.indexes, self.off)
            oracle.ionic(s.model, u_new, u, self.dense_state, self.indexes, s.dt, s.params)
            s.t += s.dt
            s.step += 1
            s.cur ^= 1
            s.launches += 1

    def device_steps(self):
        return self.launch_count()

    def launch_count(self):
        return self.sim.launches if self.sim else 0

    def keep(self, t):
        self._keep.append(t)
        return t

    def synchronize(self):
        pass


def use_oracle_engine(monkeypatch):
    """Route CardiacModel onto OracleEngine (and plain numpy host arrays instead of pinned
    ones, which need a CUDA runtime)."""
    from finitewave_b200 import model as model_mod
    oracle.build()

    def engine_for(self, tissue):
        shape = tuple(tissue.mesh.shape)
        if self._engine is None or self._engine.shape != shape:
            self._engine = OracleEngine(shape)
        return self._engine

    def alloc_host(self, shape, value):
        return np.full(shape, value, dtype=np.float64)

    monkeypatch.setattr(model_mod.CardiacModel, "_engine_for", engine_for)
    monkeypatch.setattr(model_mod.CardiacModel, "_alloc_host", alloc_host)


def host_faces_only(model):
    """Force the built-in stimuli / trackers of `model` onto their host statements."""
    for st in (model.stim_sequence.sequence if model.stim_sequence else []):
        st._native = False
    for tr in (model.tracker_sequence.sequence if model.tracker_sequence else []):
        tr._native = False
    model.async_checkpoints = False
