// host_models.cpp -- TEST-ONLY host build of finitewave_b200/csrc/models.cuh.
//
// The per-node model math of the CUDA kernels lives in a __host__ __device__
// header.  This file compiles that header with g++ so that the CPU test suite
// (-m "not gpu") can check the transcription (operation order, hoisted
// constants, read/write masks) against the oracle bit for bit, without a GPU.
// It is NOT part of the product: nothing under finitewave_b200/ builds, loads
// or calls it, and it is not a fallback path.
#include <stdint.h>

#include "../../finitewave_b200/csrc/models.cuh"

using namespace fwb;

template <int MODEL>
static void run(double *u_new, const double *u, double *const *st, int64_t n, double dt,
                const double *p)
{
    using M = Model<MODEL>;
    typename M::Consts c;
    if (!M::derive(p, dt, c)) return;
    // state accessor that also enforces the declared read / write masks
    struct IO {
        double *const *arr;
        int64_t i;
        double ld(int q) const { return (M::READ_MASK >> q) & 1 ? arr[q][i] : -12345.0; }
        void st(int q, double v) const { arr[q][i] = (M::WRITE_MASK >> q) & 1 ? v : -54321.0; }
    };
    for (int64_t i = 0; i < n; ++i) {
        IO io{st, i};
        double un = u_new[i];
        M::ionic(u[i], un, io, c);
        u_new[i] = un;
    }
}

extern "C" __attribute__((visibility("default")))
int fwb_host_ionic(int model, double *u_new, const double *u, double *const *st, int64_t n,
                   double dt, const double *p)
{
    switch (model) {
    case FWB_MODEL_ALIEV_PANFILOV: run<FWB_MODEL_ALIEV_PANFILOV>(u_new, u, st, n, dt, p); return 0;
    case FWB_MODEL_BARKLEY: run<FWB_MODEL_BARKLEY>(u_new, u, st, n, dt, p); return 0;
    case FWB_MODEL_MITCHELL_SCHAEFFER: run<FWB_MODEL_MITCHELL_SCHAEFFER>(u_new, u, st, n, dt, p); return 0;
    case FWB_MODEL_FENTON_KARMA: run<FWB_MODEL_FENTON_KARMA>(u_new, u, st, n, dt, p); return 0;
    case FWB_MODEL_LUO_RUDY91: run<FWB_MODEL_LUO_RUDY91>(u_new, u, st, n, dt, p); return 0;
    case FWB_MODEL_TP06: run<FWB_MODEL_TP06>(u_new, u, st, n, dt, p); return 0;
    case FWB_MODEL_BUENO_OROVIO: run<FWB_MODEL_BUENO_OROVIO>(u_new, u, st, n, dt, p); return 0;
    case FWB_MODEL_COURTEMANCHE: run<FWB_MODEL_COURTEMANCHE>(u_new, u, st, n, dt, p); return 0;
    }
    return -1;
}

// host build of the device's table-driven exp (finitewave_b200/csrc/fexp.cuh)
extern "C" __attribute__((visibility("default")))
void fwb_host_fexp(const double *x, double *out, int64_t n, int mode)
{
    for (int64_t i = 0; i < n; ++i)
        out[i] = mode == 0 ? fwb::fexp(x[i]) : mode == 1 ? fwb::fexp_neg(x[i]) : fwb::fexp_clamped(x[i]);
}

// host build of TP06's rearranged device path (Model<TP06>::ionic_fast) for comparison
// with the reference statement (ionic_impl<LibMath>) on the same node states
extern "C" __attribute__((visibility("default")))
int fwb_host_tp06_fast(double *u_new, const double *u, double *const *st, int64_t n, double dt,
                       const double *p)
{
    using M = Model<FWB_MODEL_TP06>;
    M::Consts c;
    if (!M::derive(p, dt, c)) return -1;
    struct IO {
        double *const *arr;
        int64_t i;
        double ld(int q) const { return (M::READ_MASK >> q) & 1 ? arr[q][i] : -12345.0; }
        void st(int q, double v) const { arr[q][i] = (M::WRITE_MASK >> q) & 1 ? v : -54321.0; }
    };
    for (int64_t i = 0; i < n; ++i) {
        IO io{st, i};
        double un = u_new[i];
        M::ionic_fast(u[i], un, io, c, io.ld(0), io.ld(3), io.ld(4));
        u_new[i] = un;
    }
    return 0;
}

// host build of the device's table-driven log (finitewave_b200/csrc/fexp.cuh)
extern "C" __attribute__((visibility("default")))
void fwb_host_flog(const double *x, double *out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) out[i] = fwb::flog(x[i]);
}

// host build of LR91's rearranged device path (Model<LR91>::ionic_fast)
extern "C" __attribute__((visibility("default")))
int fwb_host_lr91_fast(double *u_new, const double *u, double *const *st, int64_t n, double dt,
                       const double *p)
{
    using M = Model<FWB_MODEL_LUO_RUDY91>;
    M::Consts c;
    if (!M::derive(p, dt, c)) return -1;
    struct IO {
        double *const *arr;
        int64_t i;
        double ld(int q) const { return arr[q][i]; }
        void st(int q, double v) const { arr[q][i] = v; }
    };
    for (int64_t i = 0; i < n; ++i) {
        IO io{st, i};
        double un = u_new[i];
        M::ionic_fast(u[i], un, io, c, io.ld(6));
        u_new[i] = un;
    }
    return 0;
}

// host build of Courtemanche's rearranged device path (Model<COURTEMANCHE>::ionic_fast)
extern "C" __attribute__((visibility("default")))
int fwb_host_court_fast(double *u_new, const double *u, double *const *st, int64_t n, double dt,
                        const double *p)
{
    using M = Model<FWB_MODEL_COURTEMANCHE>;
    M::Consts c;
    if (!M::derive(p, dt, c)) return -1;
    struct IO {
        double *const *arr;
        int64_t i;
        double ld(int q) const { return arr[q][i]; }
        void st(int q, double v) const { arr[q][i] = v; }
    };
    for (int64_t i = 0; i < n; ++i) {
        IO io{st, i};
        double un = u_new[i];
        M::ionic_fast(u[i], un, io, c, io.ld(0), io.ld(1), io.ld(2));
        u_new[i] = un;
    }
    return 0;
}

// host build of Fenton-Karma's device path (Model<FK>::ionic_t<IO, true>: branch-free
// 1 + tanh) and of the helper itself
extern "C" __attribute__((visibility("default")))
void fwb_host_one_plus_tanh(const double *x, double *out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) out[i] = Model<FWB_MODEL_FENTON_KARMA>::one_plus_tanh_fast(x[i]);
}

extern "C" __attribute__((visibility("default")))
int fwb_host_fk_fast(double *u_new, const double *u, double *const *st, int64_t n, double dt,
                     const double *p)
{
    using M = Model<FWB_MODEL_FENTON_KARMA>;
    M::Consts c;
    if (!M::derive(p, dt, c)) return -1;
    struct IO {
        double *const *arr;
        int64_t i;
        double ld(int q) const { return arr[q][i]; }
        void st(int q, double v) const { arr[q][i] = v; }
    };
    for (int64_t i = 0; i < n; ++i) {
        IO io{st, i};
        double un = u_new[i];
        M::ionic_t<IO, true>(u[i], un, io, c);
        u_new[i] = un;
    }
    return 0;
}
