// sim.cu -- C ABI of the step backend: FwbSim and the per-step runner.
//
// fwb_sim_run reproduces, per step, the order of CardiacModel.run's loop body
// (finitewave/core/model/cardiac_model.py:164-189):
//   1. stimuli whose time has come (stim_sequence.py:65-77, stim.py:52-56)
//   2. fused diffusion + ionic step (one kernel)
//   3. native trackers whose gate passes (tracker.py:70-84) -- fused in (2)
//      for the first ActivationTime / ECG tracker, tiny extra launches otherwise
//   4. t += dt (fp64, as the reference's Python float), step += 1, swap(u, u_new)
// Host hooks (commands, state savers, user trackers) are the caller's business:
// it asks for exactly as many steps as may run before the next hook is due.
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include <cudaTypedefs.h>

#include "aux_kernels.cuh"
#include "step_kernel.cuh"

namespace fwb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

extern const ModelEntry g_entry_aliev_panfilov, g_entry_barkley, g_entry_mitchell_schaeffer,
    g_entry_fenton_karma, g_entry_luo_rudy91, g_entry_tp06, g_entry_bueno_orovio, g_entry_courtemanche, g_entry_nomodel;

const ModelEntry *model_entry(int model)
{
    switch (model) {
    case FWB_MODEL_ALIEV_PANFILOV: return &g_entry_aliev_panfilov;
    case FWB_MODEL_BARKLEY: return &g_entry_barkley;
    case FWB_MODEL_MITCHELL_SCHAEFFER: return &g_entry_mitchell_schaeffer;
    case FWB_MODEL_FENTON_KARMA: return &g_entry_fenton_karma;
    case FWB_MODEL_LUO_RUDY91: return &g_entry_luo_rudy91;
    case FWB_MODEL_TP06: return &g_entry_tp06;
    case FWB_MODEL_BUENO_OROVIO: return &g_entry_bueno_orovio;
    case FWB_MODEL_COURTEMANCHE: return &g_entry_courtemanche;
    case FWB_N_MODELS: return &g_entry_nomodel;
    default: return nullptr;
    }
}

struct Stim {
    int kind;            // 0 box, 1 node list
    int mode;
    double t, duration, value, u_max;
    int has_u_max;
    bool passed;
    StimBox box;
    const int64_t *nodes;
    int64_t n_nodes;
    std::vector<double> values;   // FWB_STIM_VOLTAGE_LIST
    int64_t fired;
};

enum { TR_ACT = 0, TR_ECG = 1, TR_POINT = 2 };
struct Tracker {
    int kind;
    double start, end;
    int64_t every;
    int64_t samples;
    // act
    double *act_t;
    double thr;
    // ecg
    const double *coords;
    int n_leads;
    double dr;
    // point
    const int64_t *items;
    const double *fill;
    int n_items;
    // ecg / point output
    double *out;
    int64_t capacity;
    bool primed;     // act: the full-grid first sample has been taken
};

}  // namespace fwb

using namespace fwb;

struct FwbSim {
    int dim, model, stencil;
    int64_t shape[3];
    Grid g;
    const ModelEntry *entry;
    const uint8_t *tissue;
    int64_t n_myo;
    double *buf[2];
    int cur;                   // buf[cur] is u
    const double *weights;
    double *state;
    double dt;
    double t;
    int64_t step;
    cudaStream_t stream;
    int device;                // the device the simulation was created on
    alignas(16) unsigned char consts[CONSTS_BYTES];
    std::vector<Stim> stims;
    std::vector<Tracker> trackers;
    double *ecg_partial;
    int64_t ecg_partial_cap;
    int64_t launches;
    int64_t device_steps;            // time steps advanced by device kernels
    const uint32_t *tile_base;
    const uint32_t *records;
    // compact-lane tile kernel: tile records, node positions, deferred-tile list, u bricks
    const uint32_t *tile_rec;
    const uint8_t *pos_of;
    uint32_t *defer;                 // [n_tiles + 2]: list, then {count, finished blocks}
    int32_t *node_of;                // [n_myo] compact index -> flat node (multi-step kernel,
                                     // packed mode of the tile kernel)
    bool packed;                     // tile kernel in packed mode (sparse tissue, no halo)
    bool small_off;                  // the multi-step cluster kernel does not fit this tissue
    bool copy_idle;                  // see StepCommon::copy_idle (fwb_sim_set_copy_idle)
    bool brick_ok;
    alignas(64) CUtensorMap tmap[2]; // of buf[0] / buf[1]
    // slab halo
    bool halo_on;
    unsigned epoch;
    uint32_t *flags;                 // local flag block
    double *peer_u[2][2];            // [side][buffer]
    int64_t peer_slices[2];
    uint32_t *peer_flags[2];
    unsigned n_side_blocks[2];
};

extern "C" const char *fwb_last_error(void) { return g_err; }
extern "C" int fwb_version(void) { return FWB_VERSION; }

extern "C" int fwb_model_n_state(int model)
{
    const ModelEntry *e = (model >= 0 && model < FWB_N_MODELS) ? model_entry(model) : nullptr;
    return e ? e->n_state : FWB_E_ARG;
}
extern "C" int fwb_model_n_params(int model)
{
    const ModelEntry *e = (model >= 0 && model < FWB_N_MODELS) ? model_entry(model) : nullptr;
    return e ? e->n_params : FWB_E_ARG;
}
extern "C" uint32_t fwb_model_read_mask(int model)
{
    const ModelEntry *e = (model >= 0 && model < FWB_N_MODELS) ? model_entry(model) : nullptr;
    return e ? e->read_mask : 0;
}
extern "C" uint32_t fwb_model_write_mask(int model)
{
    const ModelEntry *e = (model >= 0 && model < FWB_N_MODELS) ? model_entry(model) : nullptr;
    return e ? e->write_mask : 0;
}
extern "C" int fwb_stencil_k(int dim, int stencil)
{
    if (dim == 2) return stencil == FWB_STENCIL_ISO ? 5 : (stencil == FWB_STENCIL_ANISO || stencil == FWB_STENCIL_SYM) ? 9 : FWB_E_ARG;
    if (dim == 3) return stencil == FWB_STENCIL_ISO ? 7 : stencil == FWB_STENCIL_ANISO ? 19 : FWB_E_ARG;
    return FWB_E_ARG;
}

extern "C" int fwb_sim_create(FwbSim **out, int dim, const int64_t *shape, int model, int stencil,
                              const uint8_t *tissue, const uint32_t *chunk_bits,
                              const uint32_t *chunk_base, int64_t n_myo, int64_t ld,
                              const int32_t *worklist, int64_t n_work,
                              double *u, double *u_new, const double *weights, double *state,
                              const double *params, int n_params, double dt, fwb_stream_t stream)
{
    if (!out || !shape || (dim != 2 && dim != 3) || !chunk_bits || !chunk_base || !u || !u_new ||
        !weights || !tissue || !worklist || n_work < 0 || n_work % WARPS_PER_BLOCK != 0) {
        set_error("fwb_sim_create: bad argument");
        return FWB_E_ARG;
    }
    const ModelEntry *e = (model >= 0 && model < FWB_N_MODELS) ? model_entry(model) : nullptr;
    if (!e) { set_error("fwb_sim_create: unknown model %d", model); return FWB_E_ARG; }
    if (fwb_stencil_k(dim, stencil) < 0) { set_error("fwb_sim_create: unknown stencil %d", stencil); return FWB_E_ARG; }
    if (n_params != e->n_params || (n_params > 0 && !params)) {
        set_error("fwb_sim_create: model %d takes %d parameters, got %d", model, e->n_params, n_params);
        return FWB_E_ARG;
    }
    if (e->n_state > 0 && !state) { set_error("fwb_sim_create: state is NULL"); return FWB_E_ARG; }
    if (ld < n_myo) { set_error("fwb_sim_create: ld < n_myo"); return FWB_E_ARG; }
    for (int d = 0; d < dim; ++d)
        if (shape[d] < 3) { set_error("fwb_sim_create: every axis needs >= 3 nodes"); return FWB_E_ARG; }
    FwbSim *s = new FwbSim();
    s->dim = dim; s->model = model; s->stencil = stencil;
    for (int d = 0; d < 3; ++d) s->shape[d] = d < dim ? shape[d] : 1;
    s->g = make_grid(dim, shape, chunk_bits, chunk_base, ld);
    s->g.worklist = worklist; s->g.n_work = n_work;
    s->entry = e; s->tissue = tissue; s->n_myo = n_myo;
    s->buf[0] = u; s->buf[1] = u_new; s->cur = 0;
    s->weights = weights; s->state = state;
    s->dt = dt; s->t = 0.0; s->step = 0;
    s->stream = (cudaStream_t)stream;
    s->device = 0;
    cudaGetDevice(&s->device);
    memset(s->consts, 0, sizeof(s->consts));
    if (!e->derive(params, dt, s->consts)) {
        delete s;
        set_error("fwb_sim_create: model %d has a zero or non-finite divisor parameter", model);
        return FWB_E_ARG;
    }
    s->ecg_partial = nullptr; s->ecg_partial_cap = 0; s->launches = 0; s->device_steps = 0;
    s->tile_base = nullptr; s->records = nullptr;
    s->tile_rec = nullptr; s->pos_of = nullptr; s->defer = nullptr; s->brick_ok = false;
    s->node_of = nullptr; s->packed = false; s->small_off = false; s->copy_idle = false;
    s->halo_on = false; s->epoch = 0; s->flags = nullptr;
    memset(s->peer_u, 0, sizeof(s->peer_u));
    memset(s->peer_flags, 0, sizeof(s->peer_flags));
    s->peer_slices[0] = s->peer_slices[1] = 0;
    s->n_side_blocks[0] = s->n_side_blocks[1] = 0;
    *out = s;
    return 0;
}

extern "C" int fwb_sim_destroy(FwbSim *s)
{
    if (!s) return 0;
    if (s->ecg_partial) cudaFree(s->ecg_partial);
    if (s->defer) cudaFree(s->defer);
    if (s->node_of) cudaFree(s->node_of);
    delete s;
    return 0;
}

extern "C" int fwb_sim_set_time(FwbSim *s, double t, int64_t step)
{
    if (!s) return FWB_E_ARG;
    s->t = t; s->step = step;
    return 0;
}
extern "C" int fwb_sim_get_time(const FwbSim *s, double *t, int64_t *step)
{
    if (!s) return FWB_E_ARG;
    if (t) *t = s->t;
    if (step) *step = s->step;
    return 0;
}
extern "C" int fwb_sim_current_buffer(const FwbSim *s) { return s ? s->cur : FWB_E_ARG; }
extern "C" int fwb_sim_set_weights(FwbSim *s, const double *w)
{
    if (!s || !w) return FWB_E_ARG;
    s->weights = w;
    return 0;
}
extern "C" int fwb_sim_set_params(FwbSim *s, const double *params, int n_params, double dt)
{
    if (!s || n_params != s->entry->n_params) { set_error("fwb_sim_set_params: bad argument"); return FWB_E_ARG; }
    s->dt = dt;
    if (!s->entry->derive(params, dt, s->consts)) {
        set_error("fwb_sim_set_params: zero or non-finite divisor parameter");
        return FWB_E_ARG;
    }
    return 0;
}

extern "C" int fwb_sim_set_tile_base(FwbSim *s, const uint32_t *tile_base,
                                     const uint32_t *records)
{
    if (!s || ((tile_base == nullptr) != (records == nullptr))) {
        set_error("fwb_sim_set_tile_base: bad argument");
        return FWB_E_ARG;
    }
    s->tile_base = tile_base;
    s->records = records;
    return 0;
}

// tensor map of one dense u buffer for the tile kernel's brick copies: box = tile + 1-node
// halo ((2+2) x (4+2) x (32+4) doubles in 3D, (8+2) x (32+4) in 2D: a box must start on a
// 16-byte boundary, so two columns on either side); out-of-range elements of
// a box at the grid edge read as zero (they are never used: the outer ring is not tissue)
static bool make_u_tensor_map(CUtensorMap *map, const double *u, int dim, const Grid &g)
{
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            return false;
        }
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    if (dim == 3) {
        dims[0] = (cuuint64_t)g.line; dims[1] = (cuuint64_t)g.rows; dims[2] = (cuuint64_t)g.planes;
        strides[0] = (cuuint64_t)g.s_row * 8; strides[1] = (cuuint64_t)g.s_plane * 8;
        box[0] = Brick<3>::L; box[1] = Brick<3>::R; box[2] = Brick<3>::P;
    } else {
        dims[0] = (cuuint64_t)g.line; dims[1] = (cuuint64_t)g.rows;
        strides[0] = (cuuint64_t)g.s_row * 8;
        box[0] = Brick<2>::L; box[1] = Brick<2>::R;
    }
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)dim,
                              const_cast<double *>(u), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

extern "C" int fwb_sim_set_tiles(FwbSim *s, const uint32_t *tile_rec, const uint8_t *pos_of)
{
    if (!s || ((tile_rec == nullptr) != (pos_of == nullptr))) {
        set_error("fwb_sim_set_tiles: bad argument");
        return FWB_E_ARG;
    }
    s->tile_rec = tile_rec;
    s->pos_of = pos_of;
    s->brick_ok = false;
    if (!tile_rec) return 0;
    const int64_t n_tiles = s->g.n_work / WARPS_PER_BLOCK;
    if (!s->defer) {
        FWB_CUDA(cudaMalloc((void **)&s->defer, sizeof(uint32_t) * (n_tiles + 2)));
        FWB_CUDA(cudaMemsetAsync(s->defer, 0, sizeof(uint32_t) * (n_tiles + 2), s->stream));
    }
    // bricks need a line length that is a multiple of 32 nodes (16-B aligned rows, tiles that
    // are boxes) and 16-B aligned buffers; FWB_NO_BRICK=1 keeps plain loads (A/B measurements)
    const char *no = getenv("FWB_NO_BRICK");
    if ((s->g.line & 31) == 0 && !(no && no[0] == '1') &&
        ((uintptr_t)s->buf[0] & 15) == 0 && ((uintptr_t)s->buf[1] & 15) == 0)
        s->brick_ok = make_u_tensor_map(&s->tmap[0], s->buf[0], s->dim, s->g) &&
                      make_u_tensor_map(&s->tmap[1], s->buf[1], s->dim, s->g);
    return 0;
}

static int ensure_node_of(FwbSim *s);
// blocks of the step kernel = per-block ECG partial sums of a sampling step
static int64_t ecg_blocks(const FwbSim *s)
{
    const bool tile_kernel = s->entry->n_state > 4 && s->tile_rec && s->pos_of && s->defer;
    if (tile_kernel && s->packed && !s->halo_on) return (s->n_myo + BLOCK_THREADS - 1) / BLOCK_THREADS;
    return step_blocks(s->g);
}

extern "C" int fwb_sim_set_copy_idle(FwbSim *s, int on)
{
    if (!s) return FWB_E_ARG;
    s->copy_idle = on != 0;
    return 0;
}

extern "C" int fwb_sim_set_packed(FwbSim *s, int on)
{
    if (!s) return FWB_E_ARG;
    if (!on) { s->packed = false; return 0; }
    if (!s->tile_rec) { set_error("fwb_sim_set_packed: call fwb_sim_set_tiles first"); return FWB_E_STATE; }
    int rc = ensure_node_of(s);
    if (rc) return rc;
    s->packed = true;
    return 0;
}

extern "C" int fwb_sim_set_slow_offset(FwbSim *s, int64_t offset)
{
    if (!s) return FWB_E_ARG;
    s->g.slow_offset = offset;
    return 0;
}

extern "C" int fwb_sim_set_halo(FwbSim *s, uint32_t *local_flags,
                                double *peer_lo_u0, double *peer_lo_u1, int64_t peer_lo_slices,
                                uint32_t *peer_lo_flags, int64_t n_lo_blocks,
                                double *peer_hi_u0, double *peer_hi_u1, int64_t peer_hi_slices,
                                uint32_t *peer_hi_flags, int64_t n_hi_blocks)
{
    if (!s || !local_flags) { set_error("fwb_sim_set_halo: bad argument"); return FWB_E_ARG; }
    const bool lo = peer_lo_u0 != nullptr, hi = peer_hi_u0 != nullptr;
    if ((lo && (!peer_lo_u1 || !peer_lo_flags || peer_lo_slices < 3 || n_lo_blocks < 0)) ||
        (hi && (!peer_hi_u1 || !peer_hi_flags || peer_hi_slices < 3 || n_hi_blocks < 0))) {
        set_error("fwb_sim_set_halo: incomplete neighbour description");
        return FWB_E_ARG;
    }
    if (slow_stride(s->g) % 32 != 0 || slow_extent(s->g) < 4) {
        set_error("fwb_sim_set_halo: slab slices must be a multiple of 32 nodes, >= 4 slices");
        return FWB_E_UNSUPPORTED;
    }
    s->halo_on = lo || hi;
    if (s->halo_on) {
        // nothing may be loaded lazily while a neighbour's kernel spins on one of our flags
        int rc = preload_aux_kernels();
        if (!rc) rc = s->entry->preload(s->dim, s->stencil == FWB_STENCIL_SYM ? FWB_STENCIL_ANISO : s->stencil);
        if (rc) return rc;
    }
    s->flags = local_flags;
    s->peer_u[0][0] = peer_lo_u0; s->peer_u[0][1] = peer_lo_u1;
    s->peer_u[1][0] = peer_hi_u0; s->peer_u[1][1] = peer_hi_u1;
    s->peer_slices[0] = peer_lo_slices; s->peer_slices[1] = peer_hi_slices;
    s->peer_flags[0] = lo ? peer_lo_flags : nullptr;
    s->peer_flags[1] = hi ? peer_hi_flags : nullptr;
    s->n_side_blocks[0] = lo ? (unsigned)n_lo_blocks : 0;
    s->n_side_blocks[1] = hi ? (unsigned)n_hi_blocks : 0;
    s->epoch = 0;
    return 0;
}

extern "C" int fwb_sim_clear_stims(FwbSim *s)
{
    if (!s) return FWB_E_ARG;
    s->stims.clear();
    return 0;
}

static int check_mode(int mode)
{
    if (mode != FWB_STIM_VOLTAGE && mode != FWB_STIM_CURRENT && mode != FWB_STIM_VOLTAGE_LIST) {
        set_error("unknown stimulus mode %d", mode);
        return FWB_E_ARG;
    }
    return 0;
}

extern "C" int fwb_sim_add_stim_box(FwbSim *s, int mode, double t, double duration, double value,
                                    int has_u_max, double u_max, const int64_t *box)
{
    if (!s || !box) { set_error("fwb_sim_add_stim_box: bad argument"); return FWB_E_ARG; }
    if (check_mode(mode) || mode == FWB_STIM_VOLTAGE_LIST) { set_error("fwb_sim_add_stim_box: bad mode"); return FWB_E_ARG; }
    Stim st{};
    st.kind = 0; st.mode = mode; st.t = t; st.duration = duration; st.value = value;
    st.has_u_max = has_u_max; st.u_max = u_max; st.passed = false; st.fired = 0;
    int64_t lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    // (plane,row,line) <- (x,y,z) in 3D, (row,line) <- (x,y) in 2D
    const int first = 3 - s->dim;
    for (int d = 0; d < s->dim; ++d) {
        int64_t a = box[2 * d], b = box[2 * d + 1];
        const int64_t n = s->shape[d];
        if (a < 0) a = 0;
        if (b > n) b = n;
        if (b < a) b = a;
        lo[first + d] = a; hi[first + d] = b;
    }
    st.box.o0 = lo[0]; st.box.o1 = lo[1]; st.box.o2 = lo[2];
    st.box.e1 = hi[1] - lo[1]; st.box.e2 = hi[2] - lo[2];
    st.box.s0 = s->g.s_plane; st.box.s1 = s->g.s_row;
    st.box.count = (hi[0] - lo[0]) * st.box.e1 * st.box.e2;
    s->stims.push_back(st);
    return (int)s->stims.size() - 1;
}

extern "C" int fwb_sim_add_stim_nodes(FwbSim *s, int mode, double t, double duration, double value,
                                      int has_u_max, double u_max, const int64_t *nodes,
                                      int64_t n_nodes, const double *values, int64_t n_values)
{
    if (!s || (n_nodes > 0 && !nodes) || check_mode(mode)) { set_error("fwb_sim_add_stim_nodes: bad argument"); return FWB_E_ARG; }
    Stim st{};
    st.kind = 1; st.mode = mode; st.t = t; st.duration = duration; st.value = value;
    st.has_u_max = has_u_max; st.u_max = u_max; st.passed = false; st.fired = 0;
    st.nodes = nodes; st.n_nodes = n_nodes;
    if (mode == FWB_STIM_VOLTAGE_LIST) {
        if (!values || n_values <= 0) { set_error("fwb_sim_add_stim_nodes: VOLTAGE_LIST needs values"); return FWB_E_ARG; }
        st.values.assign(values, values + n_values);
    }
    s->stims.push_back(st);
    return (int)s->stims.size() - 1;
}

extern "C" int fwb_sim_stim_passed(const FwbSim *s, int id)
{
    if (!s || id < 0 || id >= (int)s->stims.size()) return FWB_E_ARG;
    return s->stims[id].passed ? 1 : 0;
}
extern "C" int fwb_sim_set_stim_passed(FwbSim *s, int id, int passed)
{
    if (!s || id < 0 || id >= (int)s->stims.size()) return FWB_E_ARG;
    s->stims[id].passed = passed != 0;
    return 0;
}

extern "C" int64_t fwb_sim_stim_fired(const FwbSim *s, int id)
{
    if (!s || id < 0 || id >= (int)s->stims.size()) return FWB_E_ARG;
    return s->stims[id].fired;
}
extern "C" int fwb_sim_set_stim_fired(FwbSim *s, int id, int64_t fired)
{
    if (!s || id < 0 || id >= (int)s->stims.size() || fired < 0) return FWB_E_ARG;
    s->stims[id].fired = fired;
    return 0;
}

extern "C" int fwb_sim_clear_trackers(FwbSim *s)
{
    if (!s) return FWB_E_ARG;
    s->trackers.clear();
    return 0;
}

extern "C" int fwb_sim_add_tracker_act(FwbSim *s, double *act_t, double threshold,
                                       double start_time, double end_time, int64_t every)
{
    if (!s || !act_t || every <= 0) { set_error("fwb_sim_add_tracker_act: bad argument"); return FWB_E_ARG; }
    Tracker tr{};
    tr.kind = TR_ACT; tr.start = start_time; tr.end = end_time; tr.every = every;
    tr.act_t = act_t; tr.thr = threshold;
    s->trackers.push_back(tr);
    return (int)s->trackers.size() - 1;
}

extern "C" int fwb_sim_add_tracker_ecg(FwbSim *s, const double *coords, int n_leads, double dr,
                                       double start_time, double end_time, int64_t every,
                                       double *out, int64_t capacity)
{
    if (!s || !coords || n_leads <= 0 || !out || every <= 0) { set_error("fwb_sim_add_tracker_ecg: bad argument"); return FWB_E_ARG; }
    for (const Tracker &o : s->trackers)
        if (o.kind == TR_ECG) { set_error("only one native ECG tracker per simulation"); return FWB_E_UNSUPPORTED; }
    Tracker tr{};
    tr.kind = TR_ECG; tr.start = start_time; tr.end = end_time; tr.every = every;
    tr.coords = coords; tr.n_leads = n_leads; tr.dr = dr; tr.out = out; tr.capacity = capacity;
    const int64_t need = step_blocks(s->g) * n_leads;
    if (need > s->ecg_partial_cap) {
        if (s->ecg_partial) cudaFree(s->ecg_partial);
        FWB_CUDA(cudaMalloc((void **)&s->ecg_partial, sizeof(double) * need));
        s->ecg_partial_cap = need;
    }
    s->trackers.push_back(tr);
    return (int)s->trackers.size() - 1;
}

extern "C" int fwb_sim_add_tracker_point(FwbSim *s, const int64_t *items, const double *fill,
                                         int n_items, double start_time, double end_time,
                                         int64_t every, double *out, int64_t capacity)
{
    if (!s || !items || !fill || n_items <= 0 || !out || every <= 0) { set_error("fwb_sim_add_tracker_point: bad argument"); return FWB_E_ARG; }
    Tracker tr{};
    tr.kind = TR_POINT; tr.start = start_time; tr.end = end_time; tr.every = every;
    tr.items = items; tr.fill = fill; tr.n_items = n_items; tr.out = out; tr.capacity = capacity;
    s->trackers.push_back(tr);
    return (int)s->trackers.size() - 1;
}

extern "C" int64_t fwb_sim_tracker_samples(const FwbSim *s, int id)
{
    if (!s || id < 0 || id >= (int)s->trackers.size()) return FWB_E_ARG;
    return s->trackers[id].samples;
}

extern "C" int64_t fwb_sim_launch_count(const FwbSim *s) { return s ? s->launches : FWB_E_ARG; }
extern "C" int64_t fwb_sim_device_steps(const FwbSim *s) { return s ? s->device_steps : FWB_E_ARG; }

static thread_local int g_step_variant = -1;
static thread_local int g_step_launches = 1;
namespace fwb {
// variants 4 / 5 (tile kernel) launch two kernels per step: the fast one and the
// reference-statement one for the tiles it hands over
void note_step_variant(int v) { g_step_variant = v; g_step_launches = (v == 4 || v == 5) ? 2 : 1; }
int last_step_launches() { return g_step_launches; }
}
extern "C" int fwb_last_step_variant(void) { return g_step_variant; }

static inline bool gate(const Tracker &tr, double t, int64_t step)
{
    // Tracker.track (core/tracker/tracker.py:70-84)
    if (tr.start > t || t > tr.end) return false;
    return step % tr.every == 0;
}

// compact index -> flat node, from the work list (one thread per work-list entry and lane)
__global__ void node_of_kernel(const int32_t *wl, int64_t n_work, const uint32_t *bits,
                               const uint32_t *base, int32_t *node_of)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_work) return;
    const int32_t c = wl[i];
    if (c < 0) return;
    const uint32_t b = bits[c];
    if ((b >> lane) & 1u) node_of[base[c] + __popc(b & ((1u << lane) - 1u))] = c * 32 + lane;
}

// Tiny tissues (README quick start: 100 x 100): how many of the next steps can run inside ONE
// launch of the multi-step cluster kernel -- no stimulus fires, nothing but (at most one,
// primed) activation-time tracker samples.  Returns 0 when the per-step path must take the
// next step.  `fused` receives that tracker, `samples` how often its gate passes.
static int64_t small_run_length(FwbSim *s, int64_t max_steps, Tracker **fused, int64_t *samples)
{
    *fused = nullptr; *samples = 0;
    if (!s->entry->launch_small || s->entry->n_state > 4 || s->halo_on || s->small_off) return 0;
    if (s->n_myo < 1) return 0;
    if (s->g.n_nodes > (int64_t)8 * SMALL_MAX_CHUNK) return 0;   // dense grid in DSMEM (8 CTAs always fit)
    const char *off = getenv("FWB_NO_SMALL_KERNEL");
    if (off && off[0] == '1') return 0;
    Tracker *act = nullptr;
    for (Tracker &tr : s->trackers) {
        if (tr.kind != TR_ACT) continue;
        if (act) return 0;                     // a second activation tracker: per-step path
        act = &tr;
    }
    double t = s->t;
    int64_t step = s->step, m = 0, hits = 0;
    while (m < max_steps) {
        for (const Stim &sm : s->stims)
            if (t >= sm.t && !sm.passed) goto done;
        for (const Tracker &tr : s->trackers) {
            if (!gate(tr, t, step)) continue;
            if (tr.kind != TR_ACT || !tr.primed) goto done;
            ++hits;
        }
        t += s->dt; ++step; ++m;
    }
done:
    *fused = act; *samples = hits;
    return m;
}

static int ensure_node_of(FwbSim *s)
{
    if (s->node_of) return 0;
    if (s->g.n_nodes >= ((int64_t)1 << 31)) { set_error("grid too large for 32-bit node ids"); return FWB_E_UNSUPPORTED; }
    FWB_CUDA(cudaMalloc((void **)&s->node_of, sizeof(int32_t) * (s->n_myo > 0 ? s->n_myo : 1)));
    const unsigned nb = (unsigned)((s->g.n_work * 32 + 255) / 256);
    if (nb) {
        node_of_kernel<<<nb, 256, 0, s->stream>>>(s->g.worklist, s->g.n_work, s->g.chunk_bits,
                                                  s->g.chunk_base, s->node_of);
        FWB_KERNEL_CHECK("node_of_kernel");
    }
    return 0;
}

static int run_small(FwbSim *s, int64_t m, Tracker *act, int64_t samples)
{
    cudaStream_t st = s->stream;
    StepCommon k;
    memset(&k, 0, sizeof(k));
    k.g = s->g; k.w = s->weights; k.state = s->state;
    SmallArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.buf[0] = s->buf[0]; sa.buf[1] = s->buf[1];
    sa.cur = s->cur; sa.n_steps = (int)m; sa.n_myo = s->n_myo;
    sa.step0 = s->step; sa.t0 = s->t;
    if (act) {
        sa.act_t = act->act_t; sa.act_thr = act->thr;
        sa.act_start = act->start; sa.act_end = act->end; sa.act_every = act->every;
    } else {
        sa.act_every = 1;
    }
    int rc = s->entry->launch_small(s->dim, s->stencil, k, s->consts, sa, st);
    if (rc) return rc;
    s->launches++;
    if (act) act->samples += samples;
    for (int64_t i = 0; i < m; ++i) s->t += s->dt;     // the reference's float accumulation
    s->step += m;
    s->device_steps += m;
    if (m & 1) s->cur ^= 1;
    return 0;
}

extern "C" int fwb_sim_run(FwbSim *s, int64_t n_steps)
{
    if (!s || n_steps < 0) { set_error("fwb_sim_run: bad argument"); return FWB_E_ARG; }
    cudaStream_t st = s->stream;
    for (int64_t it = 0; it < n_steps; ++it) {
        {
            // launch-bound sizes: as many steps as possible inside one cluster-kernel launch
            Tracker *act = nullptr;
            int64_t samples = 0;
            const int64_t m = small_run_length(s, n_steps - it > 100000 ? 100000 : n_steps - it,
                                               &act, &samples);
            if (m >= 4) {
                int rc = run_small(s, m, act, samples);
                if (rc == FWB_E_UNSUPPORTED) {
                    s->small_off = true;         // too large for one cluster: per-step kernels
                } else {
                    if (rc) return rc;
                    it += m - 1;
                    continue;
                }
            }
        }
        double *u = s->buf[s->cur];
        double *u_new = s->buf[s->cur ^ 1];
        const double t = s->t;

        // 1. stimuli (StimSequence.stimulate_next)
        bool halo_waited = false;
        for (Stim &sm : s->stims) {
            if (t >= sm.t && !sm.passed) {
                if (s->halo_on && !halo_waited) {
                    // the neighbours' stores of the previous step into our ghost slices
                    // must have landed before a stimulus edits those slices
                    int rcw = launch_halo_wait(s->peer_flags[0] ? s->flags + 0 : nullptr,
                                               s->peer_flags[1] ? s->flags + 1 : nullptr,
                                               s->epoch, st);
                    if (rcw) return rcw;
                    s->launches++;
                    halo_waited = true;
                }
                double value = sm.value;
                int mode = sm.mode;
                if (mode == FWB_STIM_VOLTAGE_LIST) {
                    if (sm.fired >= (int64_t)sm.values.size()) {
                        set_error("StimVoltageListMatrix: voltage list exhausted");
                        return FWB_E_STATE;
                    }
                    value = sm.values[sm.fired];
                    mode = FWB_STIM_VOLTAGE;
                }
                const double dt_value = s->dt * sm.value;
                int rc = sm.kind == 0
                    ? launch_stim_box(u, s->tissue, sm.box, mode, value, dt_value, sm.has_u_max, sm.u_max, st)
                    : launch_stim_nodes(u, sm.nodes, sm.n_nodes, mode, value, dt_value, sm.has_u_max, sm.u_max, st);
                if (rc) return rc;
                s->launches++;
                sm.fired++;
                sm.passed = t >= sm.t + sm.duration;   // Stim.update_status
                // nodes the solver does not update (special boundaries) change only through
                // stimuli: the activation trackers' next sample covers the whole grid again
                for (Tracker &tr : s->trackers)
                    if (tr.kind == TR_ACT) tr.primed = false;
            }
        }

        // 2+3. fused step, with the trackers whose gate passes
        StepCommon k;
        memset(&k, 0, sizeof(k));
        k.g = s->g; k.u = u; k.u_new = u_new; k.w = s->weights; k.state = s->state;
        k.t = t;
        k.tile_base = s->tile_base;
        k.records = reinterpret_cast<const uint4 *>(s->records);
        k.tile_rec = reinterpret_cast<const uint4 *>(s->tile_rec);
        k.pos_of = s->pos_of;
        if (s->defer) {
            const int64_t n_tiles = s->g.n_work / WARPS_PER_BLOCK;
            k.defer_list = s->defer;
            k.defer_ctr = s->defer + n_tiles;
        }
        if (s->packed && !s->halo_on) { k.node_of = s->node_of; k.n_packed = s->n_myo; }
        k.brick = s->brick_ok ? 1 : 0;
        k.copy_idle = s->copy_idle ? 1 : 0;
        k.tmap_host = &s->tmap[s->cur];
        if (s->halo_on) {
            Halo &h = k.halo;
            h.on = 1; h.slice = slow_stride(s->g); h.epoch = s->epoch;
            const int64_t S = slow_extent(s->g);
            const int nxt = s->cur ^ 1;
            unsigned first = 0;
            if (s->peer_flags[0]) {
                h.lo.on = 1; h.lo.first = h.slice;
                h.lo.peer_dst = s->peer_u[0][nxt] + (s->peer_slices[0] - 1) * h.slice;
                h.lo.peer_flag = s->peer_flags[0] + 1;      // we are its hi neighbour
                h.lo.flag = s->flags + 0; h.lo.counter = s->flags + 2;
                h.lo.n_blocks = s->n_side_blocks[0]; h.lo.first_block = first;
                first += h.lo.n_blocks;
            }
            if (s->peer_flags[1]) {
                h.hi.on = 1; h.hi.first = (S - 2) * h.slice;
                h.hi.peer_dst = s->peer_u[1][nxt];
                h.hi.peer_flag = s->peer_flags[1] + 0;      // we are its lo neighbour
                h.hi.flag = s->flags + 1; h.hi.counter = s->flags + 3;
                h.hi.n_blocks = s->n_side_blocks[1]; h.hi.first_block = first;
            }
        }
        bool track = false;
        Tracker *ecg = nullptr;
        bool act_fused = false;
        for (Tracker &tr : s->trackers) {
            if (!gate(tr, t, s->step)) continue;
            if (tr.kind == TR_ACT && !act_fused && tr.primed) {
                k.act_t = tr.act_t; k.act_thr = tr.thr; k.do_act = 1;
                act_fused = true; track = true; tr.samples++;
            } else if (tr.kind == TR_ECG) {
                if (tr.samples >= tr.capacity) { set_error("ECG tracker output buffer full"); return FWB_E_STATE; }
                k.do_ecg = 1; k.n_leads = tr.n_leads; k.ecg_coords = tr.coords; k.dr = tr.dr;
                k.ecg_partial = s->ecg_partial;
                ecg = &tr; track = true;
            }
        }
        int rc = s->entry->launch(s->dim, s->stencil, track, k, s->consts, st);
        if (rc) return rc;
        s->launches += last_step_launches();
        if (ecg) {
            rc = launch_ecg_finalize(s->ecg_partial, ecg_blocks(s), ecg->n_leads,
                                     ecg->out + ecg->samples * ecg->n_leads, st);
            if (rc) return rc;
            s->launches++;
            ecg->samples++;
        }
        for (Tracker &tr : s->trackers) {
            if (!gate(tr, t, s->step)) continue;
            if (tr.kind == TR_ACT) {
                if (act_fused && k.act_t == tr.act_t) continue;   // fused above
                // first sample of a tracker (and any further tracker): the whole grid,
                // including chunks without tissue that the work list skips -- their u
                // never changes afterwards, so later samples only need listed chunks
                rc = launch_act(tr.act_t, u, s->g.n_nodes, tr.thr, t, st);
                if (rc) return rc;
                s->launches++; tr.samples++; tr.primed = true;
            } else if (tr.kind == TR_POINT) {
                if (tr.samples >= tr.capacity) { set_error("point tracker output buffer full"); return FWB_E_STATE; }
                rc = launch_point_gather(tr.items, tr.fill, tr.n_items, u, s->state, s->g.ld,
                                         tr.out + tr.samples * tr.n_items, st);
                if (rc) return rc;
                s->launches++; tr.samples++;
            }
        }

        // 4. advance
        s->t += s->dt;
        s->step += 1;
        s->device_steps += 1;
        s->cur ^= 1;
        if (s->halo_on) s->epoch += 1;
    }
    return 0;
}

extern "C" int fwb_sim_halo_sync(FwbSim *s)
{
    if (!s) return FWB_E_ARG;
    if (!s->halo_on) return 0;
    // hold the stream until the neighbours' stores of the last step into our ghost
    // slices have landed (needed before the host reads or edits u)
    int rc = launch_halo_wait(s->peer_flags[0] ? s->flags + 0 : nullptr,
                              s->peer_flags[1] ? s->flags + 1 : nullptr, s->epoch, s->stream);
    if (rc) return rc;
    s->launches++;
    return 0;
}

extern "C" int fwb_diffuse(int dim, int stencil, const int64_t *shape,
                           const uint32_t *chunk_bits, const uint32_t *chunk_base, int64_t ld,
                           const int32_t *worklist, int64_t n_work,
                           const double *u, double *u_new, const double *weights,
                           fwb_stream_t stream)
{
    if (!shape || !chunk_bits || !chunk_base || !u || !u_new || !weights || !worklist ||
        n_work < 0 || n_work % WARPS_PER_BLOCK != 0 ||
        fwb_stencil_k(dim, stencil) < 0) {
        set_error("fwb_diffuse: bad argument");
        return FWB_E_ARG;
    }
    StepCommon k;
    memset(&k, 0, sizeof(k));
    k.g = make_grid(dim, shape, chunk_bits, chunk_base, ld);
    k.g.worklist = worklist; k.g.n_work = n_work;
    k.u = u; k.u_new = u_new; k.w = weights; k.state = nullptr;
    NoModel::Consts c{0.0};
    return model_entry(FWB_N_MODELS)->launch(dim, stencil, false, k, &c, (cudaStream_t)stream);
}

// ECG{2,3}DTracker.calc_ecg by hand (cpuwave2D/tracker/ecg_2d_tracker.py:61-79,
// cpuwave3D/tracker/ecg_3d_tracker.py:51-69): u_tr = W u on the updated nodes, then per lead
// sum (u_tr - u) / (d * dr) -- the fused kernel's own reduction (deterministic order),
// stand-alone.  out: [n_leads] on the device.
extern "C" int fwb_ecg(int dim, int stencil, const int64_t *shape, const uint32_t *chunk_bits,
                       const uint32_t *chunk_base, int64_t ld, const int32_t *worklist,
                       int64_t n_work, const double *u, double *u_tr, const double *weights,
                       const double *coords, int n_leads, double dr, double *out,
                       fwb_stream_t stream)
{
    if (!shape || !chunk_bits || !chunk_base || !u || !u_tr || !weights || !worklist || !coords ||
        !out || n_leads <= 0 || n_work < 0 || n_work % WARPS_PER_BLOCK != 0 ||
        fwb_stencil_k(dim, stencil) < 0) {
        set_error("fwb_ecg: bad argument");
        return FWB_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    StepCommon k;
    memset(&k, 0, sizeof(k));
    k.g = make_grid(dim, shape, chunk_bits, chunk_base, ld);
    k.g.worklist = worklist; k.g.n_work = n_work;
    k.u = u; k.u_new = u_tr; k.w = weights; k.state = nullptr;
    const int64_t blocks = step_blocks(k.g);
    if (blocks <= 0) {
        FWB_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * n_leads, st));
        return 0;
    }
    double *partial = nullptr;
    FWB_CUDA(cudaMallocAsync((void **)&partial, sizeof(double) * blocks * n_leads, st));
    k.do_ecg = 1; k.n_leads = n_leads; k.ecg_coords = coords; k.dr = dr; k.ecg_partial = partial;
    NoModel::Consts c{0.0};
    int rc = model_entry(FWB_N_MODELS)->launch(dim, stencil == FWB_STENCIL_SYM ? FWB_STENCIL_ANISO : stencil,
                                               true, k, &c, st);
    if (!rc) rc = launch_ecg_finalize(partial, blocks, n_leads, out, st);
    cudaFreeAsync(partial, st);
    return rc;
}

// Slabs of ONE process (CardiacModel.run on several GPUs, finitewave_b200/multi.py): advance all
// of them by n_steps from a single host thread.  The steps are enqueued round-robin in short
// chunks, so a slab's boundary blocks never wait for a neighbour launch that is still behind
// a full launch queue, and no host thread is ever in the middle of anything else (an
// allocation, a free) while a kernel of another slab spins on a flag.  Ends with the
// neighbours' last halo stores landed on every slab (fwb_sim_halo_sync).
extern "C" int fwb_multi_run(FwbSim **sims, int n_sims, int64_t n_steps, int64_t chunk)
{
    if (!sims || n_sims < 1 || n_steps < 0) { set_error("fwb_multi_run: bad argument"); return FWB_E_ARG; }
    if (chunk < 1) chunk = 8;
    int prev = 0;
    FWB_CUDA(cudaGetDevice(&prev));
    int rc = 0;
    for (int64_t done = 0; done < n_steps && !rc; done += chunk) {
        const int64_t k = n_steps - done < chunk ? n_steps - done : chunk;
        for (int i = 0; i < n_sims && !rc; ++i) {
            if (!sims[i]) { set_error("fwb_multi_run: NULL simulation"); rc = FWB_E_ARG; break; }
            cudaSetDevice(sims[i]->device);
            rc = fwb_sim_run(sims[i], k);
        }
    }
    for (int i = 0; i < n_sims && !rc; ++i) {
        cudaSetDevice(sims[i]->device);
        rc = fwb_sim_halo_sync(sims[i]);
    }
    cudaSetDevice(prev);
    return rc;
}
