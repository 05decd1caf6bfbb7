"""
engine.py -- device residency of one simulation (single GPU).

PyTorch is used only to own device / pinned-host buffers and to provide the
stream; every computation is a call into libfinitewave_b200.so.

Device layout: u, u_new, act_t dense; weights and state
compact SoA in myocyte order; 1 bit + 1/8 byte of index structure per node.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, shape_arr


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.FwbError("finitewave_b200 needs a CUDA device (sm_100a); "
                            "there is no CPU fallback.")


def pinned_empty(shape, dtype=torch.float64):
    return torch.empty(tuple(int(s) for s in shape), dtype=dtype, pin_memory=True)


class PeerMemory:
    """cudaMalloc'ed, zero-filled device block from fwb_dev_alloc: exportable through
    CUDA IPC (torch's caching allocator sub-allocates, which IPC handles cannot
    address).  Exposed to torch through __cuda_array_interface__."""

    def __init__(self, shape, dtype):
        self.L = lib()
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = ctypes.c_void_p(0)
        check(self.L.fwb_dev_alloc(ctypes.byref(p), max(self.nbytes, 16)), "fwb_dev_alloc")
        self.ptr = p.value
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": self.dtype.str,
                                         "data": (self.ptr, False), "version": 2}

    def tensor(self, device):
        return torch.as_tensor(self, device=device)

    def ipc_handle(self):
        buf = ctypes.create_string_buffer(self.L.fwb_ipc_handle_size())
        check(self.L.fwb_ipc_get_handle(ctypes.c_void_p(self.ptr), buf), "fwb_ipc_get_handle")
        return bytes(buf.raw)

    def __del__(self):
        try:
            if self.ptr:
                self.L.fwb_dev_free(ctypes.c_void_p(self.ptr))
                self.ptr = 0
        except Exception:
            pass


class Engine:
    """Device side of one CardiacModel instance."""

    def __init__(self, shape, device=None):
        require_cuda()
        self.L = lib()
        self.device = torch.device(device if device is not None
                                   else f"cuda:{torch.cuda.current_device()}")
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        self.n_nodes = int(np.prod(self.shape))
        self.n_chunks = (self.n_nodes + 31) // 32
        self.sim = ctypes.c_void_p(0)
        self.tissue = None       # uint8 dense, mesh == 1
        self.update = None       # uint8 dense, mesh == 1 & special == 0
        self.chunk_bits = None
        self.chunk_base = None
        self.n_myo = 0
        self.ld = 0
        self.weights = None      # [K, ld]
        self.K = 0
        self.stencil = None
        self.state = None        # [S, ld]
        self.ubuf = [None, None]
        self.staging = None
        self._offfill = None
        self.n_state = 0
        self._keep = []          # tensors borrowed by the C side (stims, trackers)

    # ---- tissue ---------------------------------------------------------
    def set_tissue(self, mesh, special_boundaries=None, halo=(False, False)):
        """mesh: host ndarray or device tensor with values 0/1/2.  halo = (lo, hi): the
        first / last slice of the slowest axis is a ghost slice owned by a neighbour
        rank (its tissue feeds weights and stimuli, its nodes are never updated here)."""
        dev = self.device
        self._halo = (bool(halo[0]), bool(halo[1]))
        self._has_special = special_boundaries is not None and bool(np.any(
            special_boundaries.cpu().numpy() if isinstance(special_boundaries, torch.Tensor)
            else np.asarray(special_boundaries)))
        self.destroy_sim()      # it borrows the index structures replaced below
        if isinstance(mesh, torch.Tensor):
            m = mesh.to(dev)
        else:
            m = torch.from_numpy(np.ascontiguousarray(mesh)).to(dev)
        tissue = (m == 1)
        # the solver never updates the outer ring (CardiacTissue.add_boundaries)
        interior = torch.zeros(self.shape, dtype=torch.bool, device=dev)
        inner = [slice(1, -1) for _ in self.shape]
        # a ghost slice is real tissue of the neighbour, not the empty outer ring
        inner[0] = slice(0 if halo[0] else 1, None if halo[1] else -1)
        interior[tuple(inner)] = True
        tissue = tissue & interior
        update = tissue
        if halo[0] or halo[1]:
            update = tissue.clone()
            if halo[0]:
                update[0] = False
            if halo[1]:
                update[-1] = False
        if special_boundaries is not None:
            sb = special_boundaries
            sb = sb.to(dev) if isinstance(sb, torch.Tensor) else torch.from_numpy(
                np.ascontiguousarray(sb)).to(dev)
            update = update & (sb == 0)     # ghost slices stay excluded
        self.tissue = tissue.to(torch.uint8).contiguous()
        self.update = update.to(torch.uint8).contiguous()
        self.chunk_bits = torch.empty(self.n_chunks, dtype=torch.int32, device=dev)
        self.chunk_base = torch.empty(self.n_chunks, dtype=torch.int32, device=dev)
        n_myo = ctypes.c_int64(0)
        check(self.L.fwb_build_chunks(_ptr(self.update), self.n_nodes, _ptr(self.chunk_bits),
                                      _ptr(self.chunk_base), ctypes.byref(n_myo), _stream()),
              "fwb_build_chunks")
        self.n_myo = int(n_myo.value)
        self.ld = max(32, (self.n_myo + 31) // 32 * 32)
        self._build_worklist(halo)

    def _build_worklist(self, halo=(False, False)):
        """Chunks with tissue in tile order (ghost slices of a slab left out)."""
        active = self.tissue
        if halo[0] or halo[1]:
            active = self.tissue.clone()
            if halo[0]:
                active[0] = 0
            if halo[1]:
                active[-1] = 0
        shp = shape_arr(self.shape)
        cap = int(self.L.fwb_worklist_capacity(self.dim, shp))
        self.worklist = torch.empty(cap, dtype=torch.int32, device=self.device)
        n_work, n_lo, n_hi = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        check(self.L.fwb_build_worklist(self.dim, shp, _ptr(active), int(halo[0]), int(halo[1]),
                                        _ptr(self.worklist), cap, ctypes.byref(n_work),
                                        ctypes.byref(n_lo), ctypes.byref(n_hi), _stream()),
              "fwb_build_worklist")
        self.n_work = int(n_work.value)
        self.halo_blocks = (int(n_lo.value), int(n_hi.value))
        self.worklist = self.worklist[:max(self.n_work, 8)].clone()
        # compact (weights / state) layout in work-list order: one contiguous range per tile
        self.tile_base = torch.zeros(self.n_work // 8 + 1, dtype=torch.int32, device=self.device)
        self.records = torch.zeros((max(self.n_work, 8), 4), dtype=torch.int32, device=self.device)
        n_myo = ctypes.c_int64(0)
        check(self.L.fwb_order_compact(_ptr(self.chunk_bits), self.n_chunks, _ptr(self.worklist),
                                       self.n_work, _ptr(self.chunk_base), _ptr(self.tile_base),
                                       _ptr(self.records), ctypes.byref(n_myo), _stream()),
              "fwb_order_compact")
        assert int(n_myo.value) == self.n_myo
        # tile view for the compact-lane tile kernel: per tile {compact base, count, origin
        # chunk, boundary}, per compact node its position inside the tile
        self.tile_rec = torch.zeros((max(self.n_work // 8, 1), 4), dtype=torch.int32,
                                    device=self.device)
        self.pos_of = torch.zeros(self.ld, dtype=torch.uint8, device=self.device)
        check(self.L.fwb_order_tiles(self.dim, shp, int(halo[0]), int(halo[1]),
                                     self.halo_blocks[0], self.halo_blocks[1],
                                     _ptr(self.chunk_bits), _ptr(self.chunk_base),
                                     _ptr(self.worklist), self.n_work, _ptr(self.tile_base),
                                     _ptr(self.tile_rec), _ptr(self.pos_of), _stream()),
              "fwb_order_tiles")

    # ---- weights --------------------------------------------------------
    def compute_weights(self, stencil, conductivity, fibers, D_al, D_ac, D_model, dt, dr):
        dev = self.device
        # the symmetric stencil only differs in its weights: the step applies them with the
        # 9-point kernel (the reference reuses diffusion_kernel_2d_aniso)
        self.stencil = _lib.STENCIL_ANISO if stencil == _lib.STENCIL_SYM else stencil
        self.K = self.L.fwb_stencil_k(self.dim, stencil)
        self.weights = torch.empty((self.K, self.ld), dtype=torch.float64, device=dev)
        cond_t, cond_s = None, 1.0
        if isinstance(conductivity, torch.Tensor):
            cond_t = conductivity.to(dev, torch.float64).contiguous()
        elif np.ndim(conductivity) == 0:
            cond_s = float(conductivity)
        else:
            cond_t = torch.from_numpy(np.ascontiguousarray(
                conductivity * np.ones(self.shape), dtype=np.float64)).to(dev)
        fib_t = None
        if fibers is not None:
            fib_t = (fibers.to(dev, torch.float64).contiguous() if isinstance(fibers, torch.Tensor)
                     else torch.from_numpy(np.ascontiguousarray(fibers, dtype=np.float64)).to(dev))
            if tuple(fib_t.shape) != (*self.shape, self.dim):
                raise ValueError(f"fibers must have shape {(*self.shape, self.dim)}")
        dr2 = dr ** 2
        check(self.L.fwb_compute_weights(
            self.dim, stencil, shape_arr(self.shape), _ptr(self.tissue), _ptr(cond_t), cond_s,
            _ptr(fib_t), float(D_al), float(D_ac), float(D_model), float(dt), float(dr2),
            _ptr(self.chunk_bits), _ptr(self.chunk_base), self.ld, _ptr(self.weights),
            _stream()), "fwb_compute_weights")
        torch.cuda.current_stream().synchronize()   # cond_t / fib_t may be freed now
        if self.sim:
            check(self.L.fwb_sim_set_weights(self.sim, _ptr(self.weights)), "fwb_sim_set_weights")

    def set_weights_dense(self, w):
        """User-supplied (*shape, K) weights (custom Stencil)."""
        w = np.ascontiguousarray(w, dtype=np.float64)
        K = w.shape[-1]
        expect = {2: (5, 9), 3: (7, 19)}[self.dim]
        if w.shape[:-1] != self.shape or K not in expect:
            raise ValueError(f"weights must have shape (*{self.shape}, K) with K in {expect}")
        self.K = K
        self.stencil = _lib.STENCIL_ISO if K in (5, 7) else _lib.STENCIL_ANISO
        d = torch.from_numpy(w).to(self.device)
        self.weights = torch.empty((K, self.ld), dtype=torch.float64, device=self.device)
        check(self.L.fwb_weights_pack(_ptr(d), _ptr(self.weights), K, self.ld, self.n_nodes,
                                      _ptr(self.chunk_bits), _ptr(self.chunk_base), _stream()),
              "fwb_weights_pack")
        torch.cuda.current_stream().synchronize()
        if self.sim:
            check(self.L.fwb_sim_set_weights(self.sim, _ptr(self.weights)), "fwb_sim_set_weights")

    def weights_dense(self):
        out = torch.empty((*self.shape, self.K), dtype=torch.float64, device=self.device)
        check(self.L.fwb_weights_unpack(_ptr(self.weights), _ptr(out), self.K, self.ld,
                                        self.n_nodes, _ptr(self.chunk_bits),
                                        _ptr(self.chunk_base), _stream()), "fwb_weights_unpack")
        return out.cpu().numpy()

    # ---- buffers --------------------------------------------------------
    def allocate(self, n_state, staging=True, peer=False):
        dev = self.device
        self.n_state = n_state
        if peer:
            # slab runs: neighbours store into these buffers through CUDA IPC mappings
            self.peer_mem = [PeerMemory(self.shape, np.float64) for _ in range(2)]
            self.peer_flags = PeerMemory((4,), np.uint32)
            self.ubuf = [m.tensor(dev) for m in self.peer_mem]
        else:
            self.ubuf = [torch.empty(self.shape, dtype=torch.float64, device=dev) for _ in range(2)]
        self.state = torch.zeros((max(n_state, 1), self.ld), dtype=torch.float64, device=dev)
        self.staging = None
        if staging:
            self._staging()

    def _staging(self):
        """Dense scratch for host <-> compact transfers (allocated on first use)."""
        if self.staging is None:
            self.staging = torch.empty(self.shape, dtype=torch.float64, device=self.device)
        return self.staging

    def upload_dense(self, which, host):
        """which: 0/1 = the two u buffers (creation order)."""
        self.ubuf[which].copy_(_as_tensor(host), non_blocking=True)

    def download_dense(self, which, host_out, rows=None):
        """rows = (r0, r1): only those slices of axis 0 (host_out has that many)."""
        src = self.ubuf[which] if rows is None else self.ubuf[which][rows[0]:rows[1]]
        _as_tensor(host_out).copy_(src, non_blocking=True)

    def update_copy_idle(self):
        """Whole-sector stores of u_new in the ring kernel (fwb_sim_set_copy_idle): allowed
        when no stimulus can touch a node the solver does not update (no special boundaries)
        and both potential buffers agree on all such nodes right now -- then storing
        u -> u_new there changes nothing.  Call after every upload of u / u_new."""
        if not self.sim:
            return
        # (FWB_COPY_IDLE=0 turns it off.  C2 on a B200: DRAM reads 1.268 -> 1.175 GB per step,
        # total traffic 1.10 x -> 1.04 x the algorithmic bytes, 46.1 -> 46.9 G upd/s in one call,
        # profiles/r2_history.md)
        ok = not getattr(self, "_has_special", False) and os.environ.get("FWB_COPY_IDLE") != "0"
        if ok:
            halo = getattr(self, "_halo", (False, False))
            own = slice(1 if halo[0] else 0, -1 if halo[1] else None)   # ghost slices are never listed
            differ = (self.ubuf[0][own] != self.ubuf[1][own]) & (self.update[own] == 0)
            ok = not bool(differ.any().item())
        check(self.L.fwb_sim_set_copy_idle(self.sim, int(ok)), "fwb_sim_set_copy_idle")

    def needs_allocation(self, n_state):
        return (self.state is None or self.ubuf[0] is None
                or self.state.shape != (max(n_state, 1), self.ld))

    def upload_state(self, slot, host, fill=None, rows=None):
        """Dense host array -> compact device row.  With `fill`, also counts (on the
        device) the non-updated nodes whose value differs from it, see off_fill().
        rows = (r0, r1): `host` only covers those slices of axis 0 (a slab's owned range);
        the other slices (ghost slices: never gathered) are set to `fill`."""
        if rows is None:
            self._staging().copy_(_as_tensor(host), non_blocking=True)
        else:
            st = self._staging()
            st.fill_(0.0 if fill is None else float(fill))
            st[rows[0]:rows[1]].copy_(_as_tensor(host), non_blocking=True)
        check(self.L.fwb_gather_compact(_ptr(self.staging), _ptr(self.state[slot]), self.n_nodes,
                                        _ptr(self.chunk_bits), _ptr(self.chunk_base), _stream()),
              "fwb_gather_compact")
        if fill is not None:
            if self._offfill is None or len(self._offfill) <= slot:
                self._offfill = torch.zeros(max(self.n_state, slot + 1), dtype=torch.int64,
                                            device=self.device)
            # (not `self._offfill[slot] = 0`: indexed assignment of a Python scalar goes through
            # a host -> device copy that waits for the stream, i.e. for the upload just queued --
            # measured: it serialised the slabs' uploads of a multi-GPU run)
            self._offfill.narrow(0, slot, 1).zero_()
            check(self.L.fwb_count_offfill(_ptr(self.staging), float(fill), self.n_nodes,
                                           _ptr(self.chunk_bits), _ptr(self._offfill[slot:]),
                                           _stream()), "fwb_count_offfill")

    def off_fill(self):
        """Per state slot: number of non-updated nodes that do not hold the model's
        init_* constant (after a StateLoader, a Command or a mesh edit).  Synchronises."""
        return [] if self._offfill is None else self._offfill.cpu().tolist()

    def download_state(self, slot, host_out, fill, keep=False, rows=None):
        """Compact device row -> dense host array.  keep=False refills the nodes the solver
        does not update with `fill` (they cannot have changed); keep=True preserves the
        values the host array holds there (one extra H2D of the array).
        rows = (r0, r1): host_out covers only those slices of axis 0."""
        if keep:
            if rows is None:
                self._staging().copy_(_as_tensor(host_out), non_blocking=True)
            else:
                self._staging()[rows[0]:rows[1]].copy_(_as_tensor(host_out), non_blocking=True)
            check(self.L.fwb_scatter_compact_keep(_ptr(self.state[slot]), _ptr(self.staging),
                                                  self.n_nodes, _ptr(self.chunk_bits),
                                                  _ptr(self.chunk_base), _stream()),
                  "fwb_scatter_compact_keep")
        else:
            check(self.L.fwb_scatter_compact(_ptr(self.state[slot]), _ptr(self._staging()),
                                             float(fill), self.n_nodes, _ptr(self.chunk_bits),
                                             _ptr(self.chunk_base), _stream()),
                  "fwb_scatter_compact")
        src = self.staging if rows is None else self.staging[rows[0]:rows[1]]
        _as_tensor(host_out).copy_(src, non_blocking=True)
        # staging is reused by the next call on the same stream: ordering is by stream

    # ---- asynchronous checkpoints ---------------------
    def snapshot_async(self, fills):
        """Snapshot u and every state row WITHOUT stalling the step loop.

        On the compute stream: one device-to-device copy of u and one compact -> dense
        scatter per state row into fresh snapshot buffers (HBM speed).  On a separate copy
        stream, after an event: the device-to-host copies into pinned buffers.  Returns
        (host_arrays, done_event): host_arrays[0] is u, [1 + slot] the state rows as dense
        numpy arrays backed by pinned memory; they are valid once done_event has completed.
        The step loop may continue as soon as this returns.  `fills[slot]` is the value of
        the nodes the solver does not update (the model's init_* constant)."""
        dev = self.device
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        snaps = [torch.empty(self.shape, dtype=torch.float64, device=dev)]
        snaps[0].copy_(self.ubuf[self.current()], non_blocking=True)
        for slot in range(self.n_state):
            d = torch.empty(self.shape, dtype=torch.float64, device=dev)
            check(self.L.fwb_scatter_compact(_ptr(self.state[slot]), _ptr(d), float(fills[slot]),
                                             self.n_nodes, _ptr(self.chunk_bits),
                                             _ptr(self.chunk_base), _stream()),
                  "fwb_scatter_compact")
            snaps.append(d)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        hosts = [torch.empty(self.shape, dtype=torch.float64, pin_memory=True) for _ in snaps]
        done = torch.cuda.Event()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            for h, d in zip(hosts, snaps):
                h.copy_(d, non_blocking=True)
                d.record_stream(self._copy_stream)   # freed only after the copy has run
            done.record(self._copy_stream)
        return [h.numpy() for h in hosts], done, hosts

    def fill_state(self, slot, value):
        self.state[slot].fill_(float(value))

    # ---- simulation object ---------------------------------------------
    def create_sim(self, model_id, params, dt, use_tma=True):
        self.destroy_sim()
        p = (ctypes.c_double * len(params))(*[float(x) for x in params])
        sim = ctypes.c_void_p(0)
        check(self.L.fwb_sim_create(
            ctypes.byref(sim), self.dim, shape_arr(self.shape), model_id, self.stencil,
            _ptr(self.tissue), _ptr(self.chunk_bits), _ptr(self.chunk_base), self.n_myo, self.ld,
            _ptr(self.worklist), self.n_work,
            _ptr(self.ubuf[0]), _ptr(self.ubuf[1]), _ptr(self.weights), _ptr(self.state),
            p, len(params), float(dt), _stream()), "fwb_sim_create")
        self.sim = sim
        self._keep = []
        if use_tma:
            check(self.L.fwb_sim_set_tile_base(sim, _ptr(self.tile_base), _ptr(self.records)),
                  "fwb_sim_set_tile_base")
            check(self.L.fwb_sim_set_tiles(sim, _ptr(self.tile_rec), _ptr(self.pos_of)),
                  "fwb_sim_set_tiles")
            # tissue that fills its spatial tiles poorly (a ventricle wall, heavy fibrosis):
            # the tile kernel packs 256 consecutive compact nodes per block instead
            fill = self.n_myo / max(1, self.n_work * 32)
            packed = os.environ.get("FWB_PACKED")
            packed = (fill < 0.85) if packed in (None, "") else packed == "1"
            if packed and not any(getattr(self, "_halo", (False, False))) \
                    and self.n_nodes < 2 ** 31:
                check(self.L.fwb_sim_set_packed(sim, 1), "fwb_sim_set_packed")

    def destroy_sim(self):
        if self.sim:
            self.L.fwb_sim_destroy(self.sim)
            self.sim = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.destroy_sim()
        except Exception:
            pass

    def set_params(self, p, dt):
        arr = (ctypes.c_double * len(p))(*[float(x) for x in p])
        check(self.L.fwb_sim_set_params(self.sim, arr, len(p), float(dt)), "fwb_sim_set_params")

    def current(self):
        return self.L.fwb_sim_current_buffer(self.sim)

    def set_time(self, t, step):
        check(self.L.fwb_sim_set_time(self.sim, float(t), int(step)), "fwb_sim_set_time")

    def get_time(self):
        t, s = ctypes.c_double(0), ctypes.c_int64(0)
        check(self.L.fwb_sim_get_time(self.sim, ctypes.byref(t), ctypes.byref(s)))
        return t.value, s.value

    def run(self, n_steps):
        check(self.L.fwb_sim_run(self.sim, int(n_steps)), "fwb_sim_run")

    def launch_count(self):
        return int(self.L.fwb_sim_launch_count(self.sim))

    def device_steps(self):
        return int(self.L.fwb_sim_device_steps(self.sim))

    def keep(self, t):
        self._keep.append(t)
        return t

    def synchronize(self):
        torch.cuda.current_stream().synchronize()

    def compact_index(self, flat):
        """compact index of flat node ids (host ints); -1 where not updated."""
        bits = self.chunk_bits.cpu().numpy().view(np.uint32)
        base = self.chunk_base.cpu().numpy().view(np.uint32)
        flat = np.asarray(flat, dtype=np.int64)
        ch, lane = flat // 32, flat % 32
        b = bits[ch]
        on = (b >> lane.astype(np.uint32)) & 1
        below = b & ((np.uint32(1) << lane.astype(np.uint32)) - np.uint32(1))
        pop = np.array([bin(int(x)).count("1") for x in below], dtype=np.int64)
        return np.where(on == 1, base[ch].astype(np.int64) + pop, -1)


def _as_tensor(a):
    if isinstance(a, torch.Tensor):
        return a
    a = np.asarray(a)
    if a.dtype != np.float64 or not a.flags.c_contiguous:
        raise ValueError("host arrays must be C-contiguous float64")
    return torch.from_numpy(a)
