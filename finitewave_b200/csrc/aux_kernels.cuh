// aux_kernels.cuh -- launchers of the small kernels used by the step runner.
#pragma once
#include "fwb_common.cuh"

namespace fwb {

// half-open index box in (plane, row, line) coordinates, already clipped
struct StimBox {
    int64_t o0, o1, o2;   // origin
    int64_t e1, e2;       // extents of row / line axes (plane extent = count / (e1*e2))
    int64_t s0, s1;       // flat strides of plane / row axes (line stride = 1)
    int64_t count;        // number of nodes in the box
};

int launch_stim_box(double *u, const uint8_t *tissue, const StimBox &b, int mode, double value,
                    double dt_value, int has_u_max, double u_max, cudaStream_t s);
int launch_stim_nodes(double *u, const int64_t *nodes, int64_t count, int mode, double value,
                      double dt_value, int has_u_max, double u_max, cudaStream_t s);
int launch_act(double *act_t, const double *u, int64_t n_nodes, double thr, double t,
               cudaStream_t s);
int launch_ecg_finalize(const double *partial, int64_t n_blocks, int n_leads, double *out,
                        cudaStream_t s);
int launch_halo_wait(const unsigned *flag_lo, const unsigned *flag_hi, unsigned epoch,
                     cudaStream_t s);
int preload_aux_kernels();
int launch_point_gather(const int64_t *items, const double *fill, int n_items, const double *u,
                        const double *state, int64_t ld, double *out, cudaStream_t s);

}  // namespace fwb
