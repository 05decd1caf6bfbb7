// weights.cu -- init-time stencil weights on the device.
//
// Replaces Stencil.compute_weights of the four on-path stencil classes:
//   IsotropicStencil2D   finitewave/cpuwave2D/stencil/isotropic_stencil_2d.py:41-69,
//                        compute_component :135-158, njit compute_weights :161-204
//   IsotropicStencil3D   finitewave/cpuwave3D/stencil/isotropic_stencil_3d.py:45-74, :112-149
//   AsymmetricStencil2D  finitewave/cpuwave2D/stencil/asymmetric_stencil_2d.py:46-84,
//                        half-step D :97-161, minor/major :204-288, njit :291-433
//   AsymmetricStencil3D  finitewave/cpuwave3D/stencil/asymmetric_stencil_3d.py:61-105, :163-458
//                        (serial njit in the reference -> one thread per node here)
//
// The half-step diffusivities (the reference's np.roll arrays) are evaluated
// on the fly; the accumulation order into every slot is the reference's, and
// all arithmetic is explicit IEEE mul/add/div (no FMA), so the weights are
// bit-identical to the reference's.  Output is compact SoA [K][ld].
#include "fwb_common.cuh"

namespace fwb {

struct WArgs {
    Grid g;
    const uint8_t *tissue;    // dense, 1 = mesh == 1
    const double *cond;       // dense or NULL
    double cond_scalar;
    const double *fibers;     // dense (*shape, dim) or NULL
    double D_al, D_ac, D_model, dt, dr2;
    double *w;                // compact SoA
};

__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

struct Ctx {
    const WArgs &A;
    int64_t n;   // centre node
    __device__ __forceinline__ int m(int64_t off) const { return A.tissue[n + off]; }
    __device__ __forceinline__ double c(int64_t node) const
    {
        return A.cond ? A.cond[node] : A.cond_scalar;
    }
};

// isotropic_stencil_2d.py:135-158: d * m0 * (m0 + (m1 == 0))
__device__ __forceinline__ double iso_component(double d, int m0, int m1)
{
    return mul(mul(d, (double)m0), (double)(m0 + (m1 == 0)));
}

// final scaling: ((w * D_model) * dt) / dr2, centre += 1
__device__ __forceinline__ double scale(const WArgs &A, double w)
{
    return dvd(mul(mul(w, A.D_model), A.dt), A.dr2);
}

template <int DIM>
__global__ void __launch_bounds__(BLOCK_THREADS) weights_iso_kernel(const __grid_constant__ WArgs A)
{
    const Grid &g = A.g;
    const int lane = threadIdx.x & 31;
    const int64_t chunk = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (chunk >= g.n_chunks) return;
    const uint32_t bits = g.chunk_bits[chunk];
    if (!((bits >> lane) & 1u)) return;
    const int64_t n = chunk * 32 + lane;
    const int64_t c = (int64_t)g.chunk_base[chunk] + __popc(bits & ((1u << lane) - 1u));
    Ctx X{A, n};
    const int64_t sP = g.s_plane, sR = g.s_row;
    const double cn = X.c(n);
    double *w = A.w + c;
    const int64_t ld = g.ld;
    if (DIM == 2) {
        // axis 0 = row stride, axis 1 = line
        const double dxx0 = mul(0.5, add(X.c(n - sR), cn)), dxx1 = mul(0.5, add(cn, X.c(n + sR)));
        const double dyy0 = mul(0.5, add(X.c(n - 1), cn)), dyy1 = mul(0.5, add(cn, X.c(n + 1)));
        const double w0 = iso_component(dxx0, X.m(-sR), X.m(sR));
        const double w1 = iso_component(dyy0, X.m(-1), X.m(1));
        const double w3 = iso_component(dyy1, X.m(1), X.m(-1));
        const double w4 = iso_component(dxx1, X.m(sR), X.m(-sR));
        const double w2 = -add(add(add(w0, w1), w3), w4);
        w[0 * ld] = scale(A, w0);
        w[1 * ld] = scale(A, w1);
        w[2 * ld] = add(scale(A, w2), 1.0);
        w[3 * ld] = scale(A, w3);
        w[4 * ld] = scale(A, w4);
    } else {
        const double dxx0 = mul(0.5, add(X.c(n - sP), cn)), dxx1 = mul(0.5, add(cn, X.c(n + sP)));
        const double dyy0 = mul(0.5, add(X.c(n - sR), cn)), dyy1 = mul(0.5, add(cn, X.c(n + sR)));
        const double dzz0 = mul(0.5, add(X.c(n - 1), cn)), dzz1 = mul(0.5, add(cn, X.c(n + 1)));
        const double w0 = iso_component(dxx0, X.m(-sP), X.m(sP));
        const double w1 = iso_component(dyy0, X.m(-sR), X.m(sR));
        const double w2 = iso_component(dzz0, X.m(-1), X.m(1));
        const double w4 = iso_component(dzz1, X.m(1), X.m(-1));
        const double w5 = iso_component(dyy1, X.m(sR), X.m(-sR));
        const double w6 = iso_component(dxx1, X.m(sP), X.m(-sP));
        const double w3 = -add(add(add(add(add(w0, w1), w2), w4), w5), w6);
        w[0 * ld] = scale(A, w0);
        w[1 * ld] = scale(A, w1);
        w[2 * ld] = scale(A, w2);
        w[3 * ld] = add(scale(A, w3), 1.0);
        w[4 * ld] = scale(A, w4);
        w[5 * ld] = scale(A, w5);
        w[6 * ld] = scale(A, w6);
    }
}

// ---- anisotropic ---------------------------------------------------------
struct Minor { double w[6]; };

// asymmetric_stencil_2d.py:204-259
__device__ __forceinline__ Minor minor_component(double d, int m0, int m1, int m2, int m3,
                                                 int m4, int m5)
{
    Minor r;
    const int m_higher = m2 + m3 + m4 + m5;
    const int m_lower = m0 + m1 + m2 + m3;
    if (m2 == 0 || m3 == 0 || m_higher < 3 || m_lower < 3) {
#pragma unroll
        for (int q = 0; q < 6; ++q) r.w[q] = 0.0;
        return r;
    }
    const double hi = (double)m_higher, lo = (double)m_lower;
    r.w[0] = dvd(mul(-d, (double)m0), lo);
    r.w[1] = dvd(mul(-d, (double)m1), lo);
    r.w[2] = mul(d, sub(dvd((double)m2, hi), dvd((double)m2, lo)));
    r.w[3] = mul(d, sub(dvd((double)m3, hi), dvd((double)m3, lo)));
    r.w[4] = dvd(mul(d, (double)m4), hi);
    r.w[5] = dvd(mul(d, (double)m5), hi);
    return r;
}

template <int DIM> struct Aniso {
    const WArgs &A;
    int64_t n;
    __device__ __forceinline__ int m(int64_t off) const { return A.tissue[n + off]; }
    // D_ab(node) * conductivity(node)   (asymmetric_stencil_2d.py:160-161, :132)
    __device__ __forceinline__ double Dc(int64_t node, int a, int b) const
    {
        const double *f = A.fibers + node * DIM;
        const double D = add(mul(A.D_ac, a == b ? 1.0 : 0.0),
                             mul(mul(sub(A.D_al, A.D_ac), f[a]), f[b]));
        return mul(D, A.cond ? A.cond[node] : A.cond_scalar);
    }
    // half-step value between `node` and `node + s` (:132-134)
    __device__ __forceinline__ double half(int64_t node, int64_t s, int a, int b) const
    {
        return mul(0.5, add(Dc(node, a, b), Dc(node + s, a, b)));
    }
};

#define ACC_SUB(slot, v) w[slot] = sub(w[slot], (v))
#define ACC_ADD(slot, v) w[slot] = add(w[slot], (v))

template <int DIM>
__global__ void __launch_bounds__(BLOCK_THREADS) weights_aniso_kernel(const __grid_constant__ WArgs A)
{
    constexpr int K = DIM == 2 ? 9 : 19;
    const Grid &g = A.g;
    const int lane = threadIdx.x & 31;
    const int64_t chunk = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (chunk >= g.n_chunks) return;
    const uint32_t bits = g.chunk_bits[chunk];
    if (!((bits >> lane) & 1u)) return;
    const int64_t n = chunk * 32 + lane;
    const int64_t c = (int64_t)g.chunk_base[chunk] + __popc(bits & ((1u << lane) - 1u));
    Aniso<DIM> X{A, n};
    double w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = 0.0;
    Minor q;

    if (DIM == 2) {
        const int64_t sI = g.s_row;
#define M2(di, dj) X.m((di) * sI + (dj))
        // q (i-1/2, j)                                 asymmetric_stencil_2d.py:336-358
        const double qx0 = mul(X.half(n - sI, sI, 0, 0), (double)M2(-1, 0));
        ACC_ADD(1, qx0); ACC_SUB(4, qx0);
        q = minor_component(X.half(n - sI, sI, 0, 1), M2(-1, -1), M2(0, -1), M2(-1, 0), M2(0, 0),
                            M2(-1, 1), M2(0, 1));
        ACC_SUB(0, q.w[0]); ACC_SUB(3, q.w[1]); ACC_SUB(1, q.w[2]);
        ACC_SUB(4, q.w[3]); ACC_SUB(2, q.w[4]); ACC_SUB(5, q.w[5]);
        // q (i, j-1/2)                                 :360-382
        const double qy0 = mul(X.half(n - 1, 1, 1, 1), (double)M2(0, -1));
        ACC_ADD(3, qy0); ACC_SUB(4, qy0);
        q = minor_component(X.half(n - 1, 1, 1, 0), M2(-1, -1), M2(-1, 0), M2(0, -1), M2(0, 0),
                            M2(1, -1), M2(1, 0));
        ACC_SUB(0, q.w[0]); ACC_SUB(1, q.w[1]); ACC_SUB(3, q.w[2]);
        ACC_SUB(4, q.w[3]); ACC_SUB(6, q.w[4]); ACC_SUB(7, q.w[5]);
        // q (i, j+1/2)                                 :384-406
        const double qy1 = mul(X.half(n, 1, 1, 1), (double)M2(0, 1));
        ACC_ADD(5, qy1); ACC_SUB(4, qy1);
        q = minor_component(X.half(n, 1, 1, 0), M2(-1, 1), M2(-1, 0), M2(0, 1), M2(0, 0),
                            M2(1, 1), M2(1, 0));
        ACC_ADD(2, q.w[0]); ACC_ADD(1, q.w[1]); ACC_ADD(5, q.w[2]);
        ACC_ADD(4, q.w[3]); ACC_ADD(8, q.w[4]); ACC_ADD(7, q.w[5]);
        // q (i+1/2, j)                                 :408-431
        const double qx1 = mul(X.half(n, sI, 0, 0), (double)M2(1, 0));
        ACC_ADD(7, qx1); ACC_SUB(4, qx1);
        q = minor_component(X.half(n, sI, 0, 1), M2(1, -1), M2(0, -1), M2(1, 0), M2(0, 0),
                            M2(1, 1), M2(0, 1));
        ACC_ADD(6, q.w[0]); ACC_ADD(3, q.w[1]); ACC_ADD(7, q.w[2]);
        ACC_ADD(4, q.w[3]); ACC_ADD(8, q.w[4]); ACC_ADD(5, q.w[5]);
#undef M2
    } else {
        const int64_t sI = g.s_plane, sJ = g.s_row;
#define M3(di, dj, dk) X.m((di) * sI + (dj) * sJ + (dk))
        double mj;
        // q (i-1/2, j, k)                              asymmetric_stencil_3d.py:211-247
        mj = mul(X.half(n - sI, sI, 0, 0), (double)M3(-1, 0, 0));
        ACC_ADD(1, mj); ACC_SUB(4, mj);
        q = minor_component(X.half(n - sI, sI, 0, 1), M3(-1, -1, 0), M3(0, -1, 0), M3(-1, 0, 0),
                            M3(0, 0, 0), M3(-1, 1, 0), M3(0, 1, 0));
        ACC_SUB(0, q.w[0]); ACC_SUB(3, q.w[1]); ACC_SUB(1, q.w[2]);
        ACC_SUB(4, q.w[3]); ACC_SUB(2, q.w[4]); ACC_SUB(5, q.w[5]);
        q = minor_component(X.half(n - sI, sI, 0, 2), M3(-1, 0, -1), M3(0, 0, -1), M3(-1, 0, 0),
                            M3(0, 0, 0), M3(-1, 0, 1), M3(0, 0, 1));
        ACC_SUB(15, q.w[0]); ACC_SUB(11, q.w[1]); ACC_SUB(1, q.w[2]);
        ACC_SUB(4, q.w[3]); ACC_SUB(17, q.w[4]); ACC_SUB(12, q.w[5]);
        // q (i+1/2, j, k)                              :249-287
        mj = mul(X.half(n, sI, 0, 0), (double)M3(1, 0, 0));
        ACC_ADD(7, mj); ACC_SUB(4, mj);
        q = minor_component(X.half(n, sI, 0, 1), M3(1, -1, 0), M3(0, -1, 0), M3(1, 0, 0),
                            M3(0, 0, 0), M3(1, 1, 0), M3(0, 1, 0));
        ACC_ADD(6, q.w[0]); ACC_ADD(3, q.w[1]); ACC_ADD(7, q.w[2]);
        ACC_ADD(4, q.w[3]); ACC_ADD(8, q.w[4]); ACC_ADD(5, q.w[5]);
        q = minor_component(X.half(n, sI, 0, 2), M3(1, 0, -1), M3(0, 0, -1), M3(1, 0, 0),
                            M3(0, 0, 0), M3(1, 0, 1), M3(0, 0, 1));
        ACC_ADD(16, q.w[0]); ACC_ADD(11, q.w[1]); ACC_ADD(7, q.w[2]);
        ACC_ADD(4, q.w[3]); ACC_ADD(18, q.w[4]); ACC_ADD(12, q.w[5]);
        // q (i, j-1/2, k)                              :289-327
        mj = mul(X.half(n - sJ, sJ, 1, 1), (double)M3(0, -1, 0));
        ACC_ADD(3, mj); ACC_SUB(4, mj);
        q = minor_component(X.half(n - sJ, sJ, 1, 0), M3(-1, -1, 0), M3(-1, 0, 0), M3(0, -1, 0),
                            M3(0, 0, 0), M3(1, -1, 0), M3(1, 0, 0));
        ACC_SUB(0, q.w[0]); ACC_SUB(1, q.w[1]); ACC_SUB(3, q.w[2]);
        ACC_SUB(4, q.w[3]); ACC_SUB(6, q.w[4]); ACC_SUB(7, q.w[5]);
        q = minor_component(X.half(n - sJ, sJ, 1, 2), M3(0, -1, -1), M3(0, 0, -1), M3(0, -1, 0),
                            M3(0, 0, 0), M3(0, -1, 1), M3(0, 0, 1));
        ACC_SUB(9, q.w[0]); ACC_SUB(11, q.w[1]); ACC_SUB(3, q.w[2]);
        ACC_SUB(4, q.w[3]); ACC_SUB(10, q.w[4]); ACC_SUB(12, q.w[5]);
        // q (i, j+1/2, k)                              :329-367
        mj = mul(X.half(n, sJ, 1, 1), (double)M3(0, 1, 0));
        ACC_ADD(5, mj); ACC_SUB(4, mj);
        q = minor_component(X.half(n, sJ, 1, 0), M3(-1, 1, 0), M3(-1, 0, 0), M3(0, 1, 0),
                            M3(0, 0, 0), M3(1, 1, 0), M3(1, 0, 0));
        ACC_ADD(2, q.w[0]); ACC_ADD(1, q.w[1]); ACC_ADD(5, q.w[2]);
        ACC_ADD(4, q.w[3]); ACC_ADD(8, q.w[4]); ACC_ADD(7, q.w[5]);
        q = minor_component(X.half(n, sJ, 1, 2), M3(0, 1, -1), M3(0, 0, -1), M3(0, 1, 0),
                            M3(0, 0, 0), M3(0, 1, 1), M3(0, 0, 1));
        ACC_ADD(13, q.w[0]); ACC_ADD(11, q.w[1]); ACC_ADD(5, q.w[2]);
        ACC_ADD(4, q.w[3]); ACC_ADD(14, q.w[4]); ACC_ADD(12, q.w[5]);
        // q (i, j, k-1/2)                              :369-407
        mj = mul(X.half(n - 1, 1, 2, 2), (double)M3(0, 0, -1));
        ACC_ADD(11, mj); ACC_SUB(4, mj);
        q = minor_component(X.half(n - 1, 1, 2, 0), M3(-1, 0, -1), M3(-1, 0, 0), M3(0, 0, -1),
                            M3(0, 0, 0), M3(1, 0, -1), M3(1, 0, 0));
        ACC_SUB(15, q.w[0]); ACC_SUB(1, q.w[1]); ACC_SUB(11, q.w[2]);
        ACC_SUB(4, q.w[3]); ACC_SUB(16, q.w[4]); ACC_SUB(7, q.w[5]);
        q = minor_component(X.half(n - 1, 1, 2, 1), M3(0, -1, -1), M3(0, -1, 0), M3(0, 0, -1),
                            M3(0, 0, 0), M3(0, 1, -1), M3(0, 1, 0));
        ACC_SUB(9, q.w[0]); ACC_SUB(3, q.w[1]); ACC_SUB(11, q.w[2]);
        ACC_SUB(4, q.w[3]); ACC_SUB(13, q.w[4]); ACC_SUB(5, q.w[5]);
        // q (i, j, k+1/2)                              :409-456
        mj = mul(X.half(n, 1, 2, 2), (double)M3(0, 0, 1));
        ACC_ADD(12, mj); ACC_SUB(4, mj);
        q = minor_component(X.half(n, 1, 2, 0), M3(-1, 0, 1), M3(-1, 0, 0), M3(0, 0, 1),
                            M3(0, 0, 0), M3(1, 0, 1), M3(1, 0, 0));
        ACC_ADD(17, q.w[0]); ACC_ADD(1, q.w[1]); ACC_ADD(12, q.w[2]);
        ACC_ADD(4, q.w[3]); ACC_ADD(18, q.w[4]); ACC_ADD(7, q.w[5]);
        q = minor_component(X.half(n, 1, 2, 1), M3(0, -1, 1), M3(0, -1, 0), M3(0, 0, 1),
                            M3(0, 0, 0), M3(0, 1, 1), M3(0, 1, 0));
        ACC_ADD(10, q.w[0]); ACC_ADD(3, q.w[1]); ACC_ADD(12, q.w[2]);
        ACC_ADD(4, q.w[3]); ACC_ADD(14, q.w[4]); ACC_ADD(5, q.w[5]);
#undef M3
    }
    double *out = A.w + c;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = scale(A, w[k]);
        if (k == 4) v = add(v, 1.0);
        out[(int64_t)k * g.ld] = v;
    }
}

// ---- symmetric 2D ---------------------------------------
// SymmetricStencil2D: cpuwave2D/stencil/symmetric_stencil_2d.py -- cell-centred tensor
// (compute_half_step_diffusion :68-118), corner weights of one cell (compute_components
// :121-151), the four cells around a node in the reference's accumulation order (njit
// compute_weights :160-250), scaling by the SCALAR D_model * dt / dr**2 (:61-62).
struct Quad { double w[4]; };

__device__ __forceinline__ Quad sym_components(double d_xx, double d_xy, double d_yx, double d_yy,
                                               int m0, int m1, int m2, int m3, double qx, double qy)
{
    Quad r;
    if (m0 + m1 + m2 + m3 < 3) {
        r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.0;
        return r;
    }
    const double qdx = add(mul(qx, d_xx), mul(qy, d_yx));
    const double qdy = add(mul(qx, d_xy), mul(qy, d_yy));
    const double w0 = sub(mul(dvd((double)(-m0), (double)(m0 + m1)), qdx),
                          mul(dvd((double)m0, (double)(m0 + m2)), qdy));
    const double w1 = add(mul(dvd((double)(-m1), (double)(m0 + m1)), qdx),
                          mul(dvd((double)m1, (double)(m1 + m3)), qdy));
    const double w2 = sub(mul(dvd((double)m2, (double)(m2 + m3)), qdx),
                          mul(dvd((double)m2, (double)(m0 + m2)), qdy));
    const double w3 = add(mul(dvd((double)m3, (double)(m2 + m3)), qdx),
                          mul(dvd((double)m3, (double)(m1 + m3)), qdy));
    r.w[0] = mul(0.5, w0); r.w[1] = mul(0.5, w1); r.w[2] = mul(0.5, w2); r.w[3] = mul(0.5, w3);
    return r;
}

__global__ void __launch_bounds__(BLOCK_THREADS) weights_sym2d_kernel(const __grid_constant__ WArgs A)
{
    const Grid &g = A.g;
    const int lane = threadIdx.x & 31;
    const int64_t chunk = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (chunk >= g.n_chunks) return;
    const uint32_t bits = g.chunk_bits[chunk];
    if (!((bits >> lane) & 1u)) return;
    const int64_t n = chunk * 32 + lane;
    const int64_t c = (int64_t)g.chunk_base[chunk] + __popc(bits & ((1u << lane) - 1u));
    const int64_t sI = g.s_row;
    Aniso<2> X{A, n};
    // tensor component (a, b) of the cell whose lower corner is `node`
    auto cell = [&](int64_t node, int a, int b) {
        return mul(0.25, add(add(add(X.Dc(node, a, b), X.Dc(node + sI, a, b)),
                                 X.Dc(node + 1, a, b)), X.Dc(node + sI + 1, a, b)));
    };
    auto comp = [&](int64_t node, double qx, double qy, int m0, int m1, int m2, int m3) {
        return sym_components(cell(node, 0, 0), cell(node, 0, 1), cell(node, 1, 0),
                              cell(node, 1, 1), m0, m1, m2, m3, qx, qy);
    };
#define M2(di, dj) X.m((di) * sI + (dj))
    double w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = 0.0;
    Quad q;
    q = comp(n - sI - 1, -1.0, -1.0, M2(-1, -1), M2(-1, 0), M2(0, -1), M2(0, 0));   // (i-1/2, j-1/2)
    ACC_ADD(0, q.w[0]); ACC_ADD(1, q.w[1]); ACC_ADD(3, q.w[2]); ACC_ADD(4, q.w[3]);
    q = comp(n - sI, -1.0, 1.0, M2(-1, 0), M2(-1, 1), M2(0, 0), M2(0, 1));          // (i-1/2, j+1/2)
    ACC_ADD(1, q.w[0]); ACC_ADD(2, q.w[1]); ACC_ADD(4, q.w[2]); ACC_ADD(5, q.w[3]);
    q = comp(n - 1, 1.0, -1.0, M2(0, -1), M2(0, 0), M2(1, -1), M2(1, 0));           // (i+1/2, j-1/2)
    ACC_ADD(3, q.w[0]); ACC_ADD(4, q.w[1]); ACC_ADD(6, q.w[2]); ACC_ADD(7, q.w[3]);
    q = comp(n, 1.0, 1.0, M2(0, 0), M2(0, 1), M2(1, 0), M2(1, 1));                  // (i+1/2, j+1/2)
    ACC_ADD(4, q.w[0]); ACC_ADD(5, q.w[1]); ACC_ADD(7, q.w[2]); ACC_ADD(8, q.w[3]);
#undef M2
    const double sc = dvd(mul(A.D_model, A.dt), A.dr2);
    double *out = A.w + c;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        double v = mul(w[k], sc);
        if (k == 4) v = add(v, 1.0);
        out[(int64_t)k * g.ld] = v;
    }
}

}  // namespace fwb

using namespace fwb;

extern "C" int fwb_compute_weights(int dim, int stencil, const int64_t *shape,
                                   const uint8_t *tissue, const double *cond, double cond_scalar,
                                   const double *fibers, double D_al, double D_ac,
                                   double D_model, double dt, double dr2,
                                   const uint32_t *chunk_bits, const uint32_t *chunk_base,
                                   int64_t ld, double *weights, fwb_stream_t stream)
{
    if ((dim != 2 && dim != 3) || !shape || !tissue || !chunk_bits || !chunk_base || !weights) {
        set_error("fwb_compute_weights: bad argument");
        return FWB_E_ARG;
    }
    if (stencil == FWB_STENCIL_SYM && dim != 2) {
        set_error("fwb_compute_weights: the symmetric stencil is 2D only");
        return FWB_E_ARG;
    }
    if ((stencil == FWB_STENCIL_ANISO || stencil == FWB_STENCIL_SYM) && !fibers) {
        set_error("Fibers must be provided for anisotropic diffusion.");
        return FWB_E_ARG;
    }
    WArgs A;
    A.g = make_grid(dim, shape, chunk_bits, chunk_base, ld);
    A.tissue = tissue; A.cond = cond; A.cond_scalar = cond_scalar; A.fibers = fibers;
    A.D_al = D_al; A.D_ac = D_ac; A.D_model = D_model; A.dt = dt; A.dr2 = dr2;
    A.w = weights;
    const unsigned blocks = (unsigned)((A.g.n_chunks + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    if (blocks == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (stencil == FWB_STENCIL_ISO) {
        if (dim == 2) weights_iso_kernel<2><<<blocks, BLOCK_THREADS, 0, s>>>(A);
        else weights_iso_kernel<3><<<blocks, BLOCK_THREADS, 0, s>>>(A);
    } else if (stencil == FWB_STENCIL_ANISO) {
        if (dim == 2) weights_aniso_kernel<2><<<blocks, BLOCK_THREADS, 0, s>>>(A);
        else weights_aniso_kernel<3><<<blocks, BLOCK_THREADS, 0, s>>>(A);
    } else if (stencil == FWB_STENCIL_SYM) {
        weights_sym2d_kernel<<<blocks, BLOCK_THREADS, 0, s>>>(A);
    } else {
        set_error("fwb_compute_weights: unknown stencil %d", stencil);
        return FWB_E_ARG;
    }
    FWB_KERNEL_CHECK("weights kernel");
    return 0;
}
