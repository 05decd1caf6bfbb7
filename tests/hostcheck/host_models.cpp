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
    M::derive(p, dt, c);
    for (int64_t i = 0; i < n; ++i) {
        double s[M::NS];
        for (int q = 0; q < M::NS; ++q) s[q] = (M::READ_MASK >> q) & 1 ? st[q][i] : -12345.0;
        double un = u_new[i];
        M::ionic(u[i], un, s, c);
        u_new[i] = un;
        for (int q = 0; q < M::NS; ++q)
            if ((M::WRITE_MASK >> q) & 1) st[q][i] = s[q];
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
    }
    return -1;
}
