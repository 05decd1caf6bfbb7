/*
 * finitewave_b200.h -- C ABI of the B200-native finitewave step backend.
 *
 * One shared library (finitewave_b200/libfinitewave_b200.so, sm_100a only).
 * Every entry point is `extern "C"`, takes plain pointers and sizes (device
 * pointers are borrowed, e.g. from torch tensors' data_ptr(); the library
 * owns only the small FwbSim bookkeeping object and its scratch), is
 * asynchronous on the caller-supplied cudaStream_t unless stated otherwise,
 * never throws, and returns
 *       0   ok
 *     > 0   a cudaError_t
 *     < 0   an argument error (FWB_E_*)
 * with a human-readable message from fwb_last_error().
 *
 * The reference (finitewave v0.8.5) has no FFI: its seam is Python template
 * methods plus function pointers to numba kernels.  Each entry point below
 * names the reference interface it replaces (paths relative to the reference
 * root, /root/reference in the build container).
 *
 * Device data layout:
 *   u, u_new, act_t        dense fp64, C order over the bounding grid (*shape)
 *   chunk_bits/chunk_base  per 32 consecutive flat nodes: bitmask of the
 *                          nodes the solver updates (mesh == 1 and
 *                          special_boundaries == 0) and the exclusive prefix
 *                          count of such nodes -> "compact index" of a node =
 *                          its position in the reference's myo_indexes array
 *   worklist               int32 ids of the chunks that contain tissue, in the
 *                          order the step kernel's blocks take them (8 per block,
 *                          -1 = padding), see fwb_build_worklist
 *   weights                compact SoA: weights[k * ld + c], k = stencil slot
 *                          in the reference's slot order, c = compact index
 *   state                  compact SoA: state[s * ld + c], s in the model's
 *                          state order (FWB_MODEL_* comments)
 */
#ifndef FINITEWAVE_B200_H
#define FINITEWAVE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FWB_VERSION 100

/* exported symbols (the library is built with -fvisibility=hidden) */
#if defined(__GNUC__)
#define FWB_API __attribute__((visibility("default")))
#else
#define FWB_API
#endif

/* error codes (< 0) */
#define FWB_E_ARG        (-1)
#define FWB_E_UNSUPPORTED (-2)
#define FWB_E_STATE      (-3)

/* models: state order = the reference's state_vars without "u";
 * parameter order = the reference's ionic_kernel argument order after dt */
#define FWB_MODEL_ALIEV_PANFILOV     0  /* state: v            params(5):  a,k,eap,mu_1,mu_2 */
#define FWB_MODEL_BARKLEY            1  /* state: v            params(3):  a,b,eap */
#define FWB_MODEL_MITCHELL_SCHAEFFER 2  /* state: h            params(5):  tau_close,tau_open,tau_in,tau_out,u_gate */
#define FWB_MODEL_FENTON_KARMA       3  /* state: v,w          params(11): tau_d,tau_o,tau_r,tau_si,tau_v_m,tau_v_p,tau_w_m,tau_w_p,k,u_c,uc_si */
#define FWB_MODEL_LUO_RUDY91         4  /* state: m,h,j,d,f,x,cai  params(15): gna,gsi,gk,gk1,gkp,gb,ko,ki,nai,nao,cao,R,T,F,PR_NaK */
#define FWB_MODEL_TP06               5  /* state: cai,casr,cass,nai,Ki,m,h,j,xr1,xr2,xs,r,s,d,f,f2,fcass,rr,oo  params(49): tp06_2d.py:204-218 order */
#define FWB_MODEL_BUENO_OROVIO       6  /* state: v,w,s        params(28): bueno_orovio_2d.py:170-178 order */
#define FWB_MODEL_COURTEMANCHE       7  /* state: nai,ki,cai,caup,carel,m,h,j_,d,f,oa,oi,ua,ui,xr,xs,fca,irel,vrel,urel,wrel  params(39): courtemanche_2d.py:157-168 order */
#define FWB_N_MODELS                 8

/* stencils; K = 5/9 (2D) or 7/19 (3D); slot order = the reference's */
#define FWB_STENCIL_ISO   0
#define FWB_STENCIL_ANISO 1
/* weights only (fwb_compute_weights, 2D): SymmetricStencil2D,
 * finitewave/cpuwave2D/stencil/symmetric_stencil_2d.py; the step applies them with the
 * 9-point kernel, i.e. as FWB_STENCIL_ANISO (the reference reuses diffusion_kernel_2d_aniso) */
#define FWB_STENCIL_SYM   2

/* stimulus modes */
#define FWB_STIM_VOLTAGE      0   /* u = value                 (StimVoltage*)  */
#define FWB_STIM_CURRENT      1   /* u += dt*value [, clamp]   (StimCurrent*)  */
#define FWB_STIM_VOLTAGE_LIST 2   /* u = values[n_fired]       (StimVoltageListMatrix3D) */

typedef void *fwb_stream_t;            /* cudaStream_t */
typedef struct FwbSim FwbSim;

FWB_API const char *fwb_last_error(void);
FWB_API int fwb_version(void);
/* number of state arrays / parameters / stencil points */
FWB_API int fwb_model_n_state(int model);
FWB_API int fwb_model_n_params(int model);
FWB_API int fwb_stencil_k(int dim, int stencil);
/* bitmask over state slots: bit s set = slot s is read / written by the step */
FWB_API uint32_t fwb_model_read_mask(int model);
FWB_API uint32_t fwb_model_write_mask(int model);

/* ------------------------------------------------------------------------
 * Tissue -> device index structures.
 * Replaces CardiacTissue.compute_myo_indexes
 * (finitewave/core/tissue/cardiac_tissue.py:54-63): instead of an int64 index
 * list (8 B/node) the device keeps 1 bit + 1/8 B per node.
 *   update_mask  dense uint8, 1 where (mesh == 1) & (special_boundaries == 0);
 *                the boundary ring must be 0 (CardiacTissue{2,3}D.add_boundaries)
 *   chunk_bits   [n_chunks]   out, n_chunks = ceil(n_nodes / 32)
 *   chunk_base   [n_chunks]   out, exclusive prefix popcount
 *   n_myo        host out (synchronises the stream)
 * ---------------------------------------------------------------------- */
FWB_API int fwb_build_chunks(const uint8_t *update_mask, int64_t n_nodes,
                     uint32_t *chunk_bits, uint32_t *chunk_base,
                     int64_t *n_myo, fwb_stream_t stream);

/* Work list of the step kernel: the chunks that contain at least one node of
 * `active` (dense uint8; normally the tissue mask mesh == 1), in tile order (8
 * consecutive entries = 2 planes x 4 rows (3D) or 8 rows (2D) of one 32-node column
 * segment), every segment padded to whole blocks with -1.  With halo_lo / halo_hi
 * the chunks of slice 1 / slice shape[0]-2 of the slowest axis (the slab's owned
 * boundary slices) come first, ghost slices 0 / shape[0]-1 are left out, and the
 * number of blocks of each boundary segment is returned.
 *   worklist   [capacity] out, capacity >= fwb_worklist_capacity(dim, shape)
 *   n_work     host out: entries used (multiple of 8)            (synchronises) */
FWB_API int64_t fwb_worklist_capacity(int dim, const int64_t *shape);
FWB_API int fwb_build_worklist(int dim, const int64_t *shape, const uint8_t *active,
                       int halo_lo, int halo_hi, int32_t *worklist, int64_t capacity,
                       int64_t *n_work, int64_t *n_lo_blocks, int64_t *n_hi_blocks,
                       fwb_stream_t stream);

/* Re-number the compact indices in WORK-LIST order (fwb_build_chunks numbers them in
 * flat order = the reference's myo_indexes order): afterwards the nodes of one tile
 * (8 consecutive work-list entries) occupy one contiguous range of every weight /
 * state row, which is what the TMA-fed step kernel copies per tile.
 *   chunk_base  [n_chunks]       rewritten
 *   tile_base   [n_work/8 + 1]   out: compact index of the first node of each tile
 *   records     [n_work][4]      out (or NULL): per work-list entry {chunk id, update
 *                                bits, compact index of the chunk's first node,
 *                                compact index of the tile's first node}; 16-B aligned
 * Every array in compact layout must be (re)filled after this call.  (synchronises) */
FWB_API int fwb_order_compact(const uint32_t *chunk_bits, int64_t n_chunks,
                      const int32_t *worklist, int64_t n_work, uint32_t *chunk_base,
                      uint32_t *tile_base, uint32_t *records, int64_t *n_myo,
                      fwb_stream_t stream);

/* Tile view of the work list for the compact-lane tile kernel (LR91 / TP06 / Courtemanche):
 * since a listed tile is a whole SPATIAL tile (8 work-list entries, -1 where a chunk holds no
 * tissue), a block can number the tile's nodes compactly (thread r <-> r-th updated node of
 * the tile: no idle lanes on sparse tissue) and fetch the tile's u neighbourhood as one brick.
 *   tile_rec  [n_work/8][4]  out: {compact index of the tile's first node, node count,
 *                            chunk id of the tile's slot 0, 1 if slab-boundary tile}
 *   pos_of    [ld] uint8     out: per compact node, (slot << 5) | lane inside its tile
 * call after fwb_order_compact with its tile_base; n_lo/n_hi_blocks from fwb_build_worklist */
FWB_API int fwb_order_tiles(int dim, const int64_t *shape, int halo_lo, int halo_hi,
                    int64_t n_lo_blocks, int64_t n_hi_blocks,
                    const uint32_t *chunk_bits, const uint32_t *chunk_base,
                    const int32_t *worklist, int64_t n_work, const uint32_t *tile_base,
                    uint32_t *tile_rec, uint8_t *pos_of, fwb_stream_t stream);

/* dense (*shape) <-> compact [ld] conversions of one array.
 * gather:  compact[c] = dense[n]            for update nodes
 * scatter: dense[n] = update ? compact[c] : fill   (all n < n_nodes)        */
FWB_API int fwb_gather_compact(const double *dense, double *compact, int64_t n_nodes,
                       const uint32_t *chunk_bits, const uint32_t *chunk_base,
                       fwb_stream_t stream);
FWB_API int fwb_scatter_compact(const double *compact, double *dense, double fill,
                        int64_t n_nodes, const uint32_t *chunk_bits,
                        const uint32_t *chunk_base, fwb_stream_t stream);
/* scatter that leaves the nodes the solver does not update untouched (their values
 * never change on the device; `dense` already holds them) */
FWB_API int fwb_scatter_compact_keep(const double *compact, double *dense, int64_t n_nodes,
                             const uint32_t *chunk_bits, const uint32_t *chunk_base,
                             fwb_stream_t stream);
/* *count (device, caller-zeroed) += number of non-updated nodes whose value differs
 * bitwise from `fill`: tells the host whether a later download may simply refill them */
FWB_API int fwb_count_offfill(const double *dense, double fill, int64_t n_nodes,
                      const uint32_t *chunk_bits, unsigned long long *count,
                      fwb_stream_t stream);
/* weights AoS dense (*shape, K) <-> compact SoA [K][ld] (user-supplied
 * Stencil results in, `model.weights` out) */
FWB_API int fwb_weights_pack(const double *dense_aos, double *compact_soa, int K, int64_t ld,
                     int64_t n_nodes, const uint32_t *chunk_bits,
                     const uint32_t *chunk_base, fwb_stream_t stream);
FWB_API int fwb_weights_unpack(const double *compact_soa, double *dense_aos, int K, int64_t ld,
                       int64_t n_nodes, const uint32_t *chunk_bits,
                       const uint32_t *chunk_base, fwb_stream_t stream);

/* ------------------------------------------------------------------------
 * Stencil weights.  Replaces Stencil.compute_weights of
 *   IsotropicStencil2D   finitewave/cpuwave2D/stencil/isotropic_stencil_2d.py:41-69,161-204
 *   IsotropicStencil3D   finitewave/cpuwave3D/stencil/isotropic_stencil_3d.py:45-74,112-149
 *   AsymmetricStencil2D  finitewave/cpuwave2D/stencil/asymmetric_stencil_2d.py:46-84,291-433
 *   AsymmetricStencil3D  finitewave/cpuwave3D/stencil/asymmetric_stencil_3d.py:61-105,163-458
 * including the D_model*dt/dr**2 scaling and the +1 on the centre slot.
 *   shape        host, dim entries
 *   tissue       dense uint8, 1 where mesh == 1 (fibrosis/empty = 0)
 *   cond         dense fp64 conductivity or NULL (then cond_scalar)
 *   fibers       dense (*shape, dim) fp64 or NULL (required for ANISO)
 *   dr2          dr**2 as evaluated by the caller
 *   weights      out, compact SoA [K][ld]
 * ---------------------------------------------------------------------- */
FWB_API int fwb_compute_weights(int dim, int stencil, const int64_t *shape,
                        const uint8_t *tissue, const double *cond, double cond_scalar,
                        const double *fibers, double D_al, double D_ac,
                        double D_model, double dt, double dr2,
                        const uint32_t *chunk_bits, const uint32_t *chunk_base,
                        int64_t ld, double *weights, fwb_stream_t stream);

/* ------------------------------------------------------------------------
 * SpiralWaveCore{2,3}DTracker on the device.  Replaces the serial njit
 * _track_tip_line / _correct_tip_pos / _apply_threshold
 * (finitewave/cpuwave2D/tracker/spiral_wave_core_2d_tracker.py:108-248; 3D: one scan per
 * slice of the last axis, cpuwave3D/tracker/spiral_wave_core_3d_tracker.py:30-48).
 *   u_prev, u   dense device potentials of the previous and the current sample
 *   out         [capacity][3] rows {x = j + dy, y = i + dx, key = (k * n_i + i) * n_j + j};
 *               rows come in no particular order: sort by key for the reference's order
 *   count       device counter, incremented once per tip (may exceed capacity: rows beyond
 *               it are dropped, the caller re-runs with a larger buffer); zero it first
 * ---------------------------------------------------------------------- */
FWB_API int fwb_tip_scan(const double *u_prev, const double *u, int dim, const int64_t *shape,
                 double threshold, double *out, unsigned capacity, unsigned *count,
                 fwb_stream_t stream);

/* ------------------------------------------------------------------------
 * Fibrosis patterns on the device.  Replaces the numpy generators
 * Diffuse{2,3}DPattern.generate / Structural{2,3}DPattern.generate
 * (finitewave/cpuwave2D/fibrosis/diffuse_2d_pattern.py:63-87, structural_2d_pattern.py:71-120,
 * finitewave/cpuwave3D/fibrosis/*.py) for meshes that only exist in device memory.
 *   mesh         int8 (*shape), device; the box is overwritten with 1 / 2
 *   box          2*dim global indices [x1, x2, y1, y2(, z1, z2)), half open
 *   block        dim block edge lengths (all 1 = diffuse); blocks are anchored at the box
 *                origin and clipped at its end
 *   density      probability of a node / block being fibrotic
 *   seed         the draw is a hash of (seed, global block id): reproducible and independent
 *                of how the tissue is cut into slabs (the reference uses numpy's global state)
 *   slow_offset  global index of local slice 0 of the slowest axis
 * ---------------------------------------------------------------------- */
FWB_API int fwb_pattern_fibrosis(int8_t *mesh, int dim, const int64_t *shape, const int64_t *box,
                         const int64_t *block, double density, unsigned long long seed,
                         int64_t slow_offset, fwb_stream_t stream);

/* ------------------------------------------------------------------------
 * LocalActivationTime{2,3}DTracker / Period{2,3}DTracker on the device.
 * Replace LocalActivationTime2DTracker.cross_threshold + _track
 * (finitewave/cpuwave2D/tracker/local_activation_time_2d_tracker.py:58-90) and
 * Period2DTracker._track (period_2d_tracker.py:38-52), which are numpy passes over model.u.
 *   fwb_lat_cross  cross = (u >= threshold) & !activated; activated updated (armed again
 *                  once u < threshold); if `layer` is given, *flag |= 1 when a crossing node
 *                  already holds a time (> -1) in that layer (=> the caller opens a new layer)
 *   fwb_lat_write  layer = where(cross, t, layer)
 *   fwb_gather_u8  out[i] = src[idx[i]]  (the detector cells of the period tracker)
 * All arrays are dense over the grid (n_nodes), device pointers; u is the pre-update
 * potential of the sampled step.
 * ---------------------------------------------------------------------- */
FWB_API int fwb_lat_cross(const double *u, int64_t n_nodes, double threshold, uint8_t *activated,
                  uint8_t *cross, const double *layer, int *flag, fwb_stream_t stream);
FWB_API int fwb_lat_write(const uint8_t *cross, int64_t n_nodes, double t, double *layer,
                  fwb_stream_t stream);
FWB_API int fwb_gather_u8(const uint8_t *src, const int64_t *idx, int64_t k, uint8_t *out,
                  fwb_stream_t stream);

/* ------------------------------------------------------------------------
 * The simulation object: the fused per-time-step path.
 * Replaces, per step, CardiacModel.run's loop body
 * (finitewave/core/model/cardiac_model.py:164-189):
 *   StimSequence.stimulate_next      core/stimulation/stim_sequence.py:65-77
 *   run_diffusion_kernel             cardiac_model.py:205-211 -> diffusion_kernel_{2d,3d}_{iso,aniso}
 *   run_ionic_kernel                 cpuwave{2D,3D}/model/<model>_{2,3}d.py ionic_kernel_{2d,3d}
 *   TrackerSequence.tracker_next     core/tracker/tracker_sequence.py:59-64 for the native trackers
 *   t += dt; step += 1; swap(u, u_new)
 * One fused kernel launch per step (plus ROI-sized stimulus launches on the
 * steps a stimulus fires and tiny finalize/gather launches on tracker
 * samples).
 * ---------------------------------------------------------------------- */
FWB_API int fwb_sim_create(FwbSim **sim, int dim, const int64_t *shape, int model, int stencil,
                   const uint8_t *tissue,              /* dense uint8, mesh == 1 (stimulus target) */
                   const uint32_t *chunk_bits, const uint32_t *chunk_base,
                   int64_t n_myo, int64_t ld,
                   const int32_t *worklist, int64_t n_work,
                   double *u, double *u_new,           /* dense; swapped every step */
                   const double *weights,              /* compact SoA [K][ld] */
                   double *state,                      /* compact SoA [S][ld] */
                   const double *params, int n_params, /* host */
                   double dt, fwb_stream_t stream);
FWB_API int fwb_sim_destroy(FwbSim *sim);

/* model.t / model.step (host scalars; t accumulates as t += dt in fp64, like
 * the reference's Python float) */
FWB_API int fwb_sim_set_time(FwbSim *sim, double t, int64_t step);
FWB_API int fwb_sim_get_time(const FwbSim *sim, double *t, int64_t *step);
/* which of the two buffers passed at creation currently is `u` (0 or 1) */
FWB_API int fwb_sim_current_buffer(const FwbSim *sim);
/* enable the TMA-fed persistent step kernel for the HBM-bound models: tile_base and
 * records from fwb_order_compact (NULL = one-block-per-tile kernel with plain loads) */
FWB_API int fwb_sim_set_tile_base(FwbSim *sim, const uint32_t *tile_base,
                          const uint32_t *records);
/* enable the compact-lane tile kernel for LR91 / TP06 / Courtemanche: tile_rec and pos_of
 * from fwb_order_tiles (NULL, NULL = off).  Also builds the tensor maps of the two u buffers
 * for the kernel's brick copies when the line length is a multiple of 32 nodes. */
FWB_API int fwb_sim_set_tiles(FwbSim *sim, const uint32_t *tile_rec, const uint8_t *pos_of);
/* tile kernel in PACKED mode (single GPU; for tissue that fills its spatial tiles poorly --
 * a ventricle wall, heavy fibrosis): a block owns 256 consecutive compact nodes instead of
 * one spatial tile, so every warp is full; u operands by plain loads.  on = 0: tile mode. */
FWB_API int fwb_sim_set_packed(FwbSim *sim, int on);
/* Ring kernel (HBM-bound models): let the lanes of a listed chunk that own no updated node
 * store u -> u_new, so that u_new is written in whole 32-byte sectors (a partially written
 * sector costs a DRAM read of the rest).  The CALLER asserts that both u buffers hold the
 * same value on every node the solver does not update (then the stores change nothing) and
 * that no stimulus can touch such a node (no special boundaries). */
FWB_API int fwb_sim_set_copy_idle(FwbSim *sim, int on);
/* rebind after the caller re-uploaded / recomputed arrays (same sizes) */
FWB_API int fwb_sim_set_weights(FwbSim *sim, const double *weights);
FWB_API int fwb_sim_set_params(FwbSim *sim, const double *params, int n_params, double dt);

/* ---- stimuli: Stim.stimulate of cpuwave{2D,3D}/stimulation/*.py ---------- */
FWB_API int fwb_sim_clear_stims(FwbSim *sim);
/* Stim{Voltage,Current}Coord{2D,3D}: half-open index box, only mesh == 1.
 * box = x1,x2,y1,y2[,z1,z2]; has_u_max = 0 -> no clamp.  Returns stim id >= 0. */
FWB_API int fwb_sim_add_stim_box(FwbSim *sim, int mode, double t, double duration, double value,
                         int has_u_max, double u_max, const int64_t *box);
/* Stim*Matrix*, StimCurrentArea*, StimVoltageListMatrix3D: flat node indices
 * (device int64, already restricted to mesh == 1 by the caller).
 * values/n_values (host) only for FWB_STIM_VOLTAGE_LIST. */
FWB_API int fwb_sim_add_stim_nodes(FwbSim *sim, int mode, double t, double duration, double value,
                           int has_u_max, double u_max, const int64_t *nodes, int64_t n_nodes,
                           const double *values, int64_t n_values);
FWB_API int fwb_sim_stim_passed(const FwbSim *sim, int stim_id);
FWB_API int fwb_sim_set_stim_passed(FwbSim *sim, int stim_id, int passed);
/* FWB_STIM_VOLTAGE_LIST: index of the next list entry (the reference keeps it in the
 * stimulus object, `self.step`, stim_voltage_list_matrix_3d.py:60-72, so it survives a
 * re-registration, a continued run and a mid-run rebuild of the simulation). */
FWB_API int64_t fwb_sim_stim_fired(const FwbSim *sim, int stim_id);
FWB_API int fwb_sim_set_stim_fired(FwbSim *sim, int stim_id, int64_t fired);

/* ---- native trackers ---------------------------------------------------- */
FWB_API int fwb_sim_clear_trackers(FwbSim *sim);
/* ActivationTime{2,3}DTracker (cpuwave2D/tracker/activation_time_2d_tracker.py:51-63),
 * gate of core/tracker/tracker.py:70-84.  act_t dense fp64, caller-initialised to -1. */
FWB_API int fwb_sim_add_tracker_act(FwbSim *sim, double *act_t, double threshold,
                            double start_time, double end_time, int64_t every);
/* ECG{2,3}DTracker (ecg_2d_tracker.py:61-79,114-152; ecg_3d_tracker.py:51-69,99-140).
 * coords device (n_leads, 3) fp64; out device [capacity][n_leads]. */
FWB_API int fwb_sim_add_tracker_ecg(FwbSim *sim, const double *coords, int n_leads, double dr,
                            double start_time, double end_time, int64_t every,
                            double *out, int64_t capacity);
/* ActionPotential / MultiVariable / Variable trackers
 * (action_potential_2d_tracker.py:46-54, multi_variable_2d_tracker.py:58-69).
 * items: n_items triples (var, flat node, compact index or -1) as device int64
 * [n_items][3]; var 0 = u, 1.. = state slot + 1; fill[n_items] device fp64 is
 * used when compact index < 0; out device [capacity][n_items]. */
FWB_API int fwb_sim_add_tracker_point(FwbSim *sim, const int64_t *items, const double *fill,
                              int n_items, double start_time, double end_time,
                              int64_t every, double *out, int64_t capacity);
/* samples taken so far by tracker id (ids in order of registration) */
FWB_API int64_t fwb_sim_tracker_samples(const FwbSim *sim, int tracker_id);

/* run exactly n_steps time steps from the current (t, step) */
FWB_API int fwb_sim_run(FwbSim *sim, int64_t n_steps);
/* kernel launches issued by fwb_sim_run so far */
FWB_API int64_t fwb_sim_launch_count(const FwbSim *sim);
/* time steps advanced by device kernels so far (one launch of the multi-step cluster kernel
 * for tiny tissues advances many) */
FWB_API int64_t fwb_sim_device_steps(const FwbSim *sim);
/* diagnostics: evaluate one of the fast-path math helpers of csrc/fexp.cuh on the device,
 * y[i] = f(x[i]); op 0 fexp, 1 fexp_fast, 2 flog, 3 frcp, 4 frcp3, 5 fsqrt (accuracy tests) */
FWB_API int fwb_devmath(int op, const double *x, double *y, int64_t n, fwb_stream_t stream);
/* diagnostics: the step-kernel variant of the calling thread's last step launch --
 * 0 one block per tile with plain loads, 1 state rows staged by TMA, 2 state + weight rows
 * staged by TMA, 3 persistent TMA ring (light models); -1 before the first step */
FWB_API int fwb_last_step_variant(void);

/* ------------------------------------------------------------------------
 * Slab decomposition across GPUs (one process per GPU; no reference counterpart:
 * the reference is single-process).
 * The tissue is cut along the slowest axis; each rank stores its owned slices
 * plus one ghost slice per neighbour (local slice 0 / shape[0]-1).  Per step the
 * blocks of the owned boundary slices store u_new both locally and directly into
 * the neighbour's ghost slice through a peer-mapped pointer (CUDA IPC over
 * NVLink) and raise a flag in the neighbour's memory; the neighbour's boundary
 * blocks of the next step wait for that flag.  No host involvement, no collective.
 *
 * fwb_dev_alloc     cudaMalloc'ed (IPC-exportable), zero-filled device memory
 * fwb_ipc_*         cudaIpcMemHandle plumbing (handle = fwb_ipc_handle_size() bytes)
 * fwb_sim_set_halo  for each side (0 = lo, 1 = hi; NULL peer_u0 = no neighbour):
 *     peer_u0/peer_u1   the neighbour's two u buffers in ITS creation order
 *     peer_slices       the neighbour's local shape[0] (its ghost slice towards us
 *                       is 0 for our hi side, peer_slices-1 for our lo side)
 *     peer_flags        the neighbour's flag block (4 uint32: [0] raised by its lo
 *                       neighbour, [1] by its hi neighbour, [2],[3] its counters)
 *     local_flags       this rank's flag block (same layout, fwb_dev_alloc(16))
 *     n_lo_blocks/n_hi_blocks from fwb_build_worklist
 *   All ranks must then call fwb_sim_run with the same n_steps, followed by
 *   fwb_sim_halo_sync before the host looks at u.
 * fwb_sim_set_slow_offset  global index of local slice 0 (ECG lead distances)
 * ---------------------------------------------------------------------- */
FWB_API int fwb_dev_alloc(void **ptr, int64_t bytes);
FWB_API int fwb_dev_free(void *ptr);
FWB_API int fwb_ipc_handle_size(void);
FWB_API int fwb_ipc_get_handle(const void *ptr, void *handle_out);
FWB_API int fwb_ipc_open_handle(const void *handle, void **ptr_out);
FWB_API int fwb_ipc_close_handle(void *ptr);
/* slabs of one process on several devices (CardiacModel.run with FWB_DEVICES): peer mapping
 * device -> peer, so that `device`'s boundary blocks can store into `peer`'s ghost slice */
FWB_API int fwb_enable_peer_access(int device, int peer);
/* advance several slab simulations of ONE process (wired with fwb_sim_set_halo through raw
 * pointers) by n_steps from a single host thread: round-robin in chunks of `chunk` steps
 * (0 = default), then fwb_sim_halo_sync on each */
FWB_API int fwb_multi_run(FwbSim **sims, int n_sims, int64_t n_steps, int64_t chunk);
FWB_API int fwb_sim_set_halo(FwbSim *sim, uint32_t *local_flags,
                     double *peer_lo_u0, double *peer_lo_u1, int64_t peer_lo_slices,
                     uint32_t *peer_lo_flags, int64_t n_lo_blocks,
                     double *peer_hi_u0, double *peer_hi_u1, int64_t peer_hi_slices,
                     uint32_t *peer_hi_flags, int64_t n_hi_blocks);
FWB_API int fwb_sim_set_slow_offset(FwbSim *sim, int64_t offset);
/* enqueue a wait for the neighbours' last ghost-slice stores (call after
 * fwb_sim_run, before the host reads u; every rank must have issued the same
 * number of steps or the wait never ends) */
FWB_API int fwb_sim_halo_sync(FwbSim *sim);

/* ------------------------------------------------------------------------
 * Single unfused pieces (used for ECG-style re-application and tests).
 * fwb_diffuse replaces diffusion_kernel_{2d,3d}_{iso,aniso}(u_new, u, w, idx).
 * ---------------------------------------------------------------------- */
FWB_API int fwb_diffuse(int dim, int stencil, const int64_t *shape,
                const uint32_t *chunk_bits, const uint32_t *chunk_base, int64_t ld,
                const int32_t *worklist, int64_t n_work,
                const double *u, double *u_new, const double *weights,
                fwb_stream_t stream);
/* ECG{2,3}DTracker.calc_ecg() by hand (cpuwave2D/tracker/ecg_2d_tracker.py:61-79 +
 * compute_ecg :114-152; cpuwave3D/tracker/ecg_3d_tracker.py:51-69, :99-140): u_tr = W u on
 * the updated nodes (dense buffer, other nodes untouched), then per lead the sum over the
 * updated nodes of (u_tr - u) / (d * dr), d = squared index distance to the lead (2D: + z^2),
 * d == 0 skipped -- the fused kernel's deterministic reduction, stand-alone.
 *   coords [n_leads][3] device, out [n_leads] device */
FWB_API int fwb_ecg(int dim, int stencil, const int64_t *shape, const uint32_t *chunk_bits,
            const uint32_t *chunk_base, int64_t ld, const int32_t *worklist, int64_t n_work,
            const double *u, double *u_tr, const double *weights, const double *coords,
            int n_leads, double dr, double *out, fwb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FINITEWAVE_B200_H */
