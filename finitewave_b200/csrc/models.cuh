// models.cuh -- per-node ionic updates of the six on-path models.
//
// Each model is a struct with
//   NS, NP            number of state arrays / reference parameters
//   READ/WRITE_MASK   which state slots the step loads / stores
//   Consts            parameters + parameter-only sub-expressions, evaluated
//                     ONCE on the host in IEEE fp64 with the reference's
//                     operation order (so the values are the reference's) and
//                     passed by value in the kernel arguments
//   derive()          host: reference parameter vector -> Consts
//   ionic()           device: one node, one time step.  `u` is the old
//                     potential, `un` enters as the diffusion result and
//                     leaves as u_new; the node's state values are read and
//                     written through `io.ld(slot)` / `io.st(slot, v)` where
//                     they are needed, so that only a few are live at a time
//                     (TP06 has 19 of them).
//   MIN_BLOCKS        __launch_bounds__ occupancy target of the step kernel
//
// The arithmetic keeps the reference's association order; the translation
// unit is compiled with -fmad=false so no multiply-add is contracted.
// Divisions by parameters and literals go through divc(): multiply by the
// correctly rounded reciprocal + one exact FMA residual correction, which
// returns the correctly rounded quotient (Markstein's theorem; checked against
// `/` on 1e9 random operands and bit-for-bit against the oracle in
// tests/test_host_models.py) in 3 FP64 instructions instead of the ~20 of the
// generic division -- whose slow path is taken whenever the numerator is zero,
// which the Heaviside-gated terms of Fenton-Karma hit on every node.
// Reference line numbers are cited per model.
#pragma once
#include <math.h>

#include "fwb_common.cuh"
#include "fexp.cuh"

#ifdef __CUDACC__
#define FWB_HD __host__ __device__ __forceinline__
#else
#define FWB_HD inline
#endif

namespace fwb {

// divisor known on the host: the value and its correctly rounded reciprocal
struct DivC {
    double c, rc;
};
inline DivC make_divc(double c)
{
    DivC d;
    d.c = c;
    d.rc = 1.0 / c;
    return d;
}
// x / d.c, correctly rounded (for finite non-zero d.c and no over/underflow)
FWB_HD double divc(double x, const DivC &d)
{
    const double q = x * d.rc;
    const double r = fma(-q, d.c, x);
    return fma(r, d.rc, q);
}

// Math policies of the two FP64-issue-bound models (LR91, TP06).  LibMath is the
// reference statement (library exp, IEEE division); it is what the host transcription
// check compiles and what the device uses for nodes whose potential is outside
// (-300, 300) mV.  FastMath is the device's normal path: table-driven exp (fexp.cuh),
// and the division fast path without its special-operand test where the denominator
// is 1 + exp(.) >= 1.
//   eu(x)  exp of an affine function of u   (|x| < 700 follows from |u| < 300)
//   en(x)  exp of a non-positive, state-dependent argument (Rush-Larsen factors)
//   ec(x)  exp of any state-dependent argument
//   dv(a, b)  a / b for b = const + exp(.) >= ~1e-2 (normal, with a normal reciprocal)
//   p15(x)    x^1.5
//   dk(x, d)  x / d.c for a host constant (literal or parameter)
struct LibMath {
    FWB_HD static double eu(double x) { return exp(x); }
    FWB_HD static double en(double x) { return exp(x); }
    FWB_HD static double ec(double x) { return exp(x); }
    FWB_HD static double dv(double a, double b) { return a / b; }
    FWB_HD static double p15(double x) { return pow(x, 1.5); }
    FWB_HD static double dk(double x, const DivC &d) { return divc(x, d); }
    static constexpr bool FUSE_RATE = false;
};
struct FastMath {
    FWB_HD static double eu(double x) { return fexp(x); }
    FWB_HD static double en(double x) { return fexp_neg(x); }
    FWB_HD static double ec(double x) { return fexp_clamped(x); }
    FWB_HD static double dv(double a, double b)
    {
        const double y = frcp(b);
        if (a == 1.0) return y;               // folded at compile time for literal numerators
        const double q = a * y;
        return fma(fma(-b, q, a), y, q);
    }
    FWB_HD static double p15(double x) { return x * sqrt(x); }   // x^1.5, x > 0
    // x / constant as one multiplication by the rounded reciprocal (<= 1 ulp off the
    // quotient; the exact 3-instruction divc() stays on the LibMath path)
    FWB_HD static double dk(double x, const DivC &d) { return x * d.rc; }
    static constexpr bool FUSE_RATE = true;
};
constexpr double FAST_MATH_U_LIMIT = 300.0;

// x / k for a literal k (the reciprocal is folded at compile time)
#define FWB_DIVK(x, k) ::fwb::divc((x), ::fwb::DivC{(k), 1.0 / (k)})
// the same through a math policy E (LR91 / TP06 / Courtemanche): exact on LibMath, one
// multiplication on FastMath
#define FWB_EDIVK(x, k) E::dk((x), ::fwb::DivC{(k), 1.0 / (k)})

// every DivC of a Consts block must be usable (parameter finite and non-zero)
inline bool divc_ok(const DivC &d) { return d.rc == d.rc && d.rc != 0.0 && (d.rc - d.rc) == 0.0; }

template <int MODEL> struct Model;

// ---------------------------------------------------------------------------
// Aliev-Panfilov -- finitewave/cpuwave2D/model/aliev_panfilov_2d.py:133-170
// (calc_v), :173-202 (ionic_kernel_2d); 3D: cpuwave3D/model/aliev_panfilov_3d.py:53-84
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_ALIEV_PANFILOV> {
    static constexpr int NS = 1, NP = 5, MIN_BLOCKS = 4;
    static constexpr bool USE_TMA = true;   // step_kernel_tma for HBM-bound models
    static constexpr uint32_t READ_MASK = 0x1, WRITE_MASK = 0x1;
    struct Consts { double dt, a, k, eap, mu_1, mu_2; };
    static bool derive(const double *p, double dt, Consts &c)
    {
        c.dt = dt; c.a = p[0]; c.k = p[1]; c.eap = p[2]; c.mu_1 = p[3]; c.mu_2 = p[4];
        return true;
    }
    template <class IO>
    FWB_HD static void ionic(double u, double &un, IO &io, const Consts &c)
    {
        double v = io.ld(0);
        v += (-c.dt * (c.eap + (c.mu_1 * v) / (c.mu_2 + u)) * (v + c.k * u * (u - c.a - 1.)));
        io.st(0, v);
        un += c.dt * (-c.k * u * (u - c.a) * (u - 1.) - u * v);
    }
};

// ---------------------------------------------------------------------------
// Barkley -- cpuwave2D/model/barkley_2d.py:113-137, :140-166; barkley_3d.py:53-83
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_BARKLEY> {
    static constexpr int NS = 1, NP = 3, MIN_BLOCKS = 4;
    static constexpr bool USE_TMA = true;   // step_kernel_tma for HBM-bound models
    static constexpr uint32_t READ_MASK = 0x1, WRITE_MASK = 0x1;
    struct Consts { double dt, b; DivC a, eap; };
    static bool derive(const double *p, double dt, Consts &c)
    {
        c.dt = dt; c.a = make_divc(p[0]); c.b = p[1]; c.eap = make_divc(p[2]);
        return divc_ok(c.a) && divc_ok(c.eap);
    }
    template <class IO>
    FWB_HD static void ionic(double u, double &un, IO &io, const Consts &c)
    {
        double v = io.ld(0);
        v += c.dt * (u - v);
        io.st(0, v);
        un += divc(c.dt * (u * (1 - u) * (u - divc(v + c.b, c.a))), c.eap);
    }
};

// ---------------------------------------------------------------------------
// Mitchell-Schaeffer -- cpuwave2D/model/mitchell_schaeffer_2d.py:102-186,
// :188-217; mitchell_schaeffer_3d.py:57-95.  strict `u < u_gate`.
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_MITCHELL_SCHAEFFER> {
    static constexpr int NS = 1, NP = 5, MIN_BLOCKS = 4;
    static constexpr bool USE_TMA = true;   // step_kernel_tma for HBM-bound models
    static constexpr uint32_t READ_MASK = 0x1, WRITE_MASK = 0x1;
    struct Consts { double dt, u_gate; DivC tau_close, tau_open, tau_in, tau_out; };
    static bool derive(const double *p, double dt, Consts &c)
    {
        c.dt = dt; c.tau_close = make_divc(p[0]); c.tau_open = make_divc(p[1]);
        c.tau_in = make_divc(p[2]); c.tau_out = make_divc(p[3]); c.u_gate = p[4];
        return divc_ok(c.tau_close) && divc_ok(c.tau_open) && divc_ok(c.tau_in) &&
               divc_ok(c.tau_out);
    }
    template <class IO>
    FWB_HD static void ionic(double u, double &un, IO &io, const Consts &c)
    {
        double h = io.ld(0);
        h += (u < c.u_gate) ? divc(c.dt * (1.0 - h), c.tau_open) : divc(c.dt * (-h), c.tau_close);
        io.st(0, h);
        const double C = (u * u) * (1 - u);
        const double J_in = divc(h * C, c.tau_in);
        const double J_out = divc(-u, c.tau_out);
        un += c.dt * (J_in + J_out);
    }
};

// ---------------------------------------------------------------------------
// Fenton-Karma -- cpuwave2D/model/fenton_karma_2d.py:143-292, :294-328;
// fenton_karma_3d.py:61-97.  Both Heavisides are 1 at u == u_c; J_si is
// evaluated with v (not w) exactly as the reference does (:326).
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_FENTON_KARMA> {
    static constexpr int NS = 2, NP = 11, MIN_BLOCKS = 4;
    static constexpr bool USE_TMA = true;   // step_kernel_tma for HBM-bound models
    static constexpr uint32_t READ_MASK = 0x3, WRITE_MASK = 0x3;
    struct Consts {
        double dt, k, u_c, uc_si;
        DivC tau_d, tau_o, tau_r, tau_v_m, tau_v_p, tau_w_m, tau_w_p, two_tau_si;
    };
    static bool derive(const double *p, double dt, Consts &c)
    {
        c.dt = dt; c.tau_d = make_divc(p[0]); c.tau_o = make_divc(p[1]);
        c.tau_r = make_divc(p[2]); c.two_tau_si = make_divc(2 * p[3]);
        c.tau_v_m = make_divc(p[4]); c.tau_v_p = make_divc(p[5]);
        c.tau_w_m = make_divc(p[6]); c.tau_w_p = make_divc(p[7]);
        c.k = p[8]; c.u_c = p[9]; c.uc_si = p[10];
        return divc_ok(c.tau_d) && divc_ok(c.tau_o) && divc_ok(c.tau_r) &&
               divc_ok(c.two_tau_si) && divc_ok(c.tau_v_m) && divc_ok(c.tau_v_p) &&
               divc_ok(c.tau_w_m) && divc_ok(c.tau_w_p);
    }
    // 1 + tanh(x) as 2 g / (1 + g), g = exp(2x): CUDA's tanh switches algorithm at
    // |x| = 0.55, and on an excited sheet k (u - uc_si) straddles that value inside most
    // warps of the plateau -- both paths get executed (C2: 0.905 -> 0.85 of HBM once half of
    // the tissue is excited).  This form costs the same on every lane (table-driven fexp +
    // frcp3, fexp.cuh) and is accurate to a few ulp RELATIVE for any x, where 1 + tanh(x)
    // cancels for x << 0.  |x| >= 300 and NaN take the library statement.
    FWB_HD static double one_plus_tanh_fast(double x)
    {
        if (fabs(x) < 300.0) {
            const double g = fexp_fast(2.0 * x);
            return (2.0 * g) * frcp3(1.0 + g);
        }
        return 1 + tanh(x);
    }
    template <class IO>
    FWB_HD static void ionic(double u, double &un, IO &io, const Consts &c)
    {
#ifdef __CUDA_ARCH__
        ionic_t<IO, true>(u, un, io, c);
#else
        ionic_t<IO, false>(u, un, io, c);     // the reference statement (host check)
#endif
    }
    template <class IO, bool FAST>
    FWB_HD static void ionic_t(double u, double &un, IO &io, const Consts &c)
    {
        const double H1 = (c.u_c - u >= 0) ? 1.0 : 0.0;
        const double H2 = (u - c.u_c >= 0) ? 1.0 : 0.0;
        double v = io.ld(0), w = io.ld(1);
        v += c.dt * (divc(H1 * (1 - v), c.tau_v_m) - divc(H2 * v, c.tau_v_p));
        w += c.dt * (divc(H1 * (1 - w), c.tau_w_m) - divc(H2 * w, c.tau_w_p));
        io.st(0, v);
        io.st(1, w);
        const double J_fi = divc(-(v * H2 * (1 - u) * (u - c.u_c)), c.tau_d);
        const double J_so = divc(u * H1, c.tau_o) + divc(H2, c.tau_r);
        const double x = c.k * (u - c.uc_si);
        const double th1 = FAST ? one_plus_tanh_fast(x) : 1 + tanh(x);
        const double J_si = divc(-v * th1, c.two_tau_si);
        un += c.dt * (-J_fi - J_so - J_si);
    }
};

// ---------------------------------------------------------------------------
// Bueno-Orovio minimal ventricular model -- cpuwave2D/model/bueno_orovio_2d.py:183-433
// (point functions), :435-475 (ionic_kernel_2d); bueno_orovio_3d.py.  SURVEY 8f row f1.
// state: v, w, s.  Time constants chosen by a threshold are pre-inverted pairs.
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_BUENO_OROVIO> {
    static constexpr int NS = 3, NP = 28, MIN_BLOCKS = 4;
    static constexpr bool USE_TMA = true;
    static constexpr uint32_t READ_MASK = 0x7, WRITE_MASK = 0x7;
    struct Consts {
        double dt, u_o, u_u, theta_v, theta_w, theta_v_m, theta_o;
        double tau_w1_m, tau_w21_m, k_w_m, u_w_m, tau_so1, tau_so21, k_so, u_so, k_s, u_s, w_inf_;
        DivC tau_v1_m, tau_v2_m, tau_v_p, tau_w_p, tau_fi, tau_o1, tau_o2, tau_s1, tau_s2,
            tau_si, tau_w_inf;
    };
    static bool derive(const double *p, double dt, Consts &c)
    {
        c.dt = dt; c.u_o = p[0]; c.u_u = p[1]; c.theta_v = p[2]; c.theta_w = p[3];
        c.theta_v_m = p[4]; c.theta_o = p[5];
        c.tau_v1_m = make_divc(p[6]); c.tau_v2_m = make_divc(p[7]); c.tau_v_p = make_divc(p[8]);
        c.tau_w1_m = p[9]; c.tau_w21_m = p[10] - p[9]; c.k_w_m = p[11]; c.u_w_m = p[12];
        c.tau_w_p = make_divc(p[13]); c.tau_fi = make_divc(p[14]);
        c.tau_o1 = make_divc(p[15]); c.tau_o2 = make_divc(p[16]);
        c.tau_so1 = p[17]; c.tau_so21 = p[18] - p[17]; c.k_so = p[19]; c.u_so = p[20];
        c.tau_s1 = make_divc(p[21]); c.tau_s2 = make_divc(p[22]); c.k_s = p[23]; c.u_s = p[24];
        c.tau_si = make_divc(p[25]); c.tau_w_inf = make_divc(p[26]); c.w_inf_ = p[27];
        return divc_ok(c.tau_v1_m) && divc_ok(c.tau_v2_m) && divc_ok(c.tau_v_p) &&
               divc_ok(c.tau_w_p) && divc_ok(c.tau_fi) && divc_ok(c.tau_o1) && divc_ok(c.tau_o2) &&
               divc_ok(c.tau_s1) && divc_ok(c.tau_s2) && divc_ok(c.tau_si) && divc_ok(c.tau_w_inf);
    }
    template <class IO>
    FWB_HD static void ionic(double u, double &un, IO &io, const Consts &c)
    {
        const double dt = c.dt;
        // calc_v :183-213 (calc_v_inf :411-421, calc_tau_v_m :330-345)
        const double v_inf = (u < c.theta_v_m) ? 1.0 : 0.0;
        double v = io.ld(0);
        {
            const double num = v_inf - v;
            const double below = ((u - c.theta_v_m) < 0) ? divc(num, c.tau_v1_m) : divc(num, c.tau_v2_m);
            v += dt * (((u - c.theta_v) < 0) ? below : divc(-v, c.tau_v_p));
        }
        io.st(0, v);
        // calc_w :216-246 (calc_w_inf :424-433, calc_tau_w_m :348-362)
        const double w_inf = ((u - c.theta_o) < 0) ? 1 - divc(u, c.tau_w_inf) : c.w_inf_;
        const double tau_w_m = c.tau_w1_m + c.tau_w21_m * (1 + tanh(c.k_w_m * (u - c.u_w_m))) * 0.5;
        double w = io.ld(1);
        w += dt * (((u - c.theta_w) < 0) ? (w_inf - w) / tau_w_m : divc(-w, c.tau_w_p));
        io.st(1, w);
        // calc_s :249-270 (calc_tau_s :381-393)
        double s = io.ld(2);
        {
            const double num = dt * ((1 + tanh(c.k_s * (u - c.u_s))) * 0.5 - s);
            s += ((u - c.theta_w) < 0) ? divc(num, c.tau_s1) : divc(num, c.tau_s2);
        }
        io.st(2, s);
        // calc_Jfi :273-289, calc_Jso :292-310 (calc_tau_o :396-408, calc_tau_so :365-378),
        // calc_Jsi :313-327
        const double Hv = ((u - c.theta_v) >= 0) ? 1.0 : 0.0;
        const double Hw = ((u - c.theta_w) >= 0) ? 1.0 : 0.0;
        const double J_fi = divc(-v * Hv * (u - c.theta_v) * (c.u_u - u), c.tau_fi);
        const double tau_so = c.tau_so1 + c.tau_so21 * (1 + tanh(c.k_so * (u - c.u_so))) * 0.5;
        const double o_num = (u - c.u_o) * (1 - Hw);
        const double J_so = (((u - c.theta_o) < 0) ? divc(o_num, c.tau_o1) : divc(o_num, c.tau_o2)) +
                            Hw / tau_so;
        const double J_si = divc(-Hw * w * s, c.tau_si);
        un += dt * (-J_fi - J_so - J_si);
    }
};

// ---------------------------------------------------------------------------
// Luo-Rudy 1991 -- cpuwave2D/model/luo_rudy91_2d.py:158-443, :446-510;
// luo_rudy91_3d.py:63-132.  Gates are forward Euler (calc_gating_var :158-182),
// I_Na uses the new m,h,j, I_si the old d,f, I_K the old x.
// state: m,h,j,d,f,x,cai
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_LUO_RUDY91> {
#ifndef FWB_LR91_MIN_BLOCKS
#define FWB_LR91_MIN_BLOCKS 4
#endif
    static constexpr int NS = 7, NP = 15, MIN_BLOCKS = FWB_LR91_MIN_BLOCKS;
    static constexpr bool USE_TMA = false;   // step_kernel_tma for HBM-bound models
    static constexpr uint32_t READ_MASK = 0x7f, WRITE_MASK = 0x7f;
    struct Consts {
        double dt, gna, gsi, gkp, gb;
        double E_Na, E_K, G_K, E_K1, G_K1;   // parameter-only (:478, :337-338, :496, :404)
        // fast path only (ionic_fast): exp(offset * slope) of the shared-slope exponentials
        double k47, k32, kd, kf1, kf2, kf3, kx1, kx2, kxa, kxb, kdn, kfn;
        double ec[8];      // fexp's reduction / polynomial constants (fexp_fill_consts)
    };
    static bool derive(const double *p, double dt, Consts &c)
    {
        fexp_fill_consts(c.ec);
        const double gk = p[2], gk1 = p[3], ko = p[6], ki = p[7], nai = p[8], nao = p[9],
                     R = p[11], T = p[12], F = p[13], PR_NaK = p[14];
        c.dt = dt; c.gna = p[0]; c.gsi = p[1]; c.gkp = p[4]; c.gb = p[5];
        c.E_Na = (R * T / F) * log(nao / nai);
        c.E_K = (R * T / F) * log((ko + PR_NaK * nao) / (ki + PR_NaK * nai));
        c.G_K = gk * sqrt(ko / 5.4);
        c.E_K1 = (R * T / F) * log(ko / ki);
        c.G_K1 = gk1 * sqrt(ko / 5.4);
        c.k47 = exp(-0.1 * 47.13); c.k32 = exp(-0.1 * 32.); c.kd = exp(0.05 * 44.);
        c.kf1 = exp(-0.02 * 30.); c.kf2 = exp(-0.2 * 30.); c.kf3 = exp(0.15 * 28.);
        c.kx1 = exp(-0.06 * 20.); c.kx2 = exp(-0.04 * 20.);
        c.kxa = exp(0.04 * (77. - 35.)); c.kxb = exp(-0.04 * 35.);
        c.kdn = 0.095 * exp(0.01 * 5.); c.kfn = 0.012 * exp(-0.008 * 28.);
        return true;
    }
    FWB_HD static double gate(double var, double dt, double alpha, double beta)
    {
        const double tau = 1. / (alpha + beta);
        const double inf = alpha / (alpha + beta);
        var += dt * (inf - var) / tau;
        return var;
    }
    static constexpr bool HAS_FAST = true;
    // the rearranged path needs |u| < 300 mV and a positive, normal cai (flog); NaNs fail
    // the comparisons and take the reference statement
    template <class IO> FWB_HD static bool fast_ok(double u, const IO &io, const Consts &)
    {
        const double cai = io.ld(6);
        return fabs(u) < FAST_MATH_U_LIMIT && cai > 1e-300 && cai < 1e300;
    }
    template <class IO>
    FWB_HD static void ionic_fastpath(double u, double &un, IO &io, const Consts &c)
    {
        ionic_fast(u, un, io, c, io.ld(6));
    }
    template <class IO>
    FWB_HD static void ionic_ref(double u, double &un, IO &io, const Consts &c)
    {
        ionic_impl<IO, LibMath>(u, un, io, c, io.ld(6));
    }
    template <class IO>
    FWB_HD static void ionic(double u, double &un, IO &io, const Consts &c)
    {
#ifdef __CUDA_ARCH__
        if (fast_ok(u, io, c)) ionic_fastpath(u, un, io, c);
        else
#endif
            ionic_ref(u, un, io, c);
    }

    // ------------------------------------------------------------------------------------
    // The device's normal path: the equations of ionic_impl rearranged for the FP64 pipe
    // (identities only; tests/test_host_models.py compares the two on random node states):
    //  * forward-Euler gate: with tau = 1 / (a + b), inf = a / (a + b) the reference's
    //    dt (inf - x) / tau is dt (a - x (a + b)): no division at all
    //  * a = n1 / A, b = n2 / B are put over the common denominator A B: one reciprocal
    //    (frcp3) per gate; K_1x = a / (a + b) likewise
    //  * exponentials with commensurable slopes share one evaluation: exp(-0.002 u) is the
    //    common root of -0.008 / -0.01 / -0.02 / -0.04 / -0.06 (six multiplications replace
    //    four evaluations; the 30th power carries 30 x the root's half-ulp error), -0.1 / -0.2
    //    share another, 0.05 / 0.15 a third; Xi needs none:
    //    (exp(.04 (u + 77)) - 1) / exp(.04 (u + 35)) = exp(1.68) - exp(-1.4) exp(-.04 u)
    //  * exp(-2.535e-7 u), |u| < 300: four Taylor terms (remainder < 2e-18)
    //  * log(cai) by flog
    // ------------------------------------------------------------------------------------
    FWB_HD static double gatef(double x, double dt, double a, double ab) { return fma(dt, fma(-x, ab, a), x); }
    template <class IO>
    FWB_HD static void ionic_fast(double u, double &un, IO &io, const Consts &c, double cai)
    {
        const double dt = c.dt;
        const double *ec = c.ec;       // kernel-parameter bank -> uniform registers
        auto EX = [ec](double x) { return fexp_fast_p(x, ec); };
        auto EXN = [ec](double x) { return fexp_fast_p(x < -700.0 ? -700.0 : x, ec); };
        auto EXC = [ec](double x) {
            x = x < -700.0 ? -700.0 : x;
            return fexp_fast_p(x > 700.0 ? 700.0 : x, ec);
        };
        (void)EXN; (void)EXC;
        const double r1 = EX(-0.002 * u), r2 = r1 * r1, r4 = r2 * r2;  // exp(-0.008 u)
        const double g001 = r4 * r1;                                   // exp(-0.01 u)
        const double g02 = g001 * g001, g04 = g02 * g02, g06 = g04 * g02;
        // exp(-0.1 u) keeps its own evaluation: 1 - k47 g01 cancels around u = -47.13, where a
        // 50th power's 5e-15 would show as 3e-11 in alpha_m
        const double g01 = EX(-0.1 * u), g20 = g01 * g01;               // exp(-0.2 u)
        const double g05 = EX(0.05 * u), g15 = g05 * g05 * g05;
        // ---- I_Na (calc_ina :185-241)
        double ah, sh, aj, sj;                                          // alpha, alpha + beta
        if (u >= -40.) {
            ah = 0.;
            sh = frcp3(0.13 * (1. + EX((u + 10.66) * (-1. / 11.1))));
            aj = 0.;
            const double z = -2.535e-07 * u;          // |z| < 7.7e-5 (fast_ok: |u| < 300)
            const double ez = fma(z, fma(z, fma(z, 1. / 6., 0.5), 1.), 1.);
            sj = 0.3 * ez * frcp3(fma(c.k32, g01, 1.));
        } else {
            ah = 0.135 * EX((80. + u) * (-1. / 6.8));
            sh = ah + (3.56 * EX(0.079 * u) + 3.1e5 * EX(0.35 * u));
            const double a = 1. + EX(0.311 * (u + 79.23));
            const double b = 1. + EX(-0.1378 * (u + 40.14));
            const double na = (-1.2714e5 * EX(0.2444 * u) -
                               3.474e-5 * EX(-0.04391 * u)) * (u + 37.78);
            const double nb = 0.1212 * EX(-0.01052 * u);
            const double r = frcp3(a * b);
            aj = na * b * r;
            sj = fma(na, b, nb * a) * r;
        }
        const double am = 0.32 * (u + 47.13) * frcp3(fma(-c.k47, g01, 1.));
        const double sm = am + 0.08 * EX(u * (-1. / 11.));
        const double m = gatef(io.ld(0), dt, am, sm);
        const double h = gatef(io.ld(1), dt, ah, sh);
        const double j = gatef(io.ld(2), dt, aj, sj);
        io.st(0, m); io.st(1, h); io.st(2, j);
        const double ina = c.gna * m * m * m * h * j * (u - c.E_Na);
        // ---- I_si (calc_isk :244-294): old d, f
        double d = io.ld(3), f = io.ld(4);
        const double E_Si = fma(-13.0287, flog(cai), 7.7);
        const double I_Si = c.gsi * d * f * (u - E_Si);
        {
            const double n1 = c.kdn * g001;
            const double a = 1. + EX(-0.072 * (u - 5.));
            const double n2 = 0.07 * EX(-0.017 * (u + 44.));
            const double b = fma(c.kd, g05, 1.);
            const double r = frcp3(a * b);
            d = gatef(d, dt, n1 * b * r, fma(n1, b, n2 * a) * r);
        }
        {
            const double n1 = c.kfn * r4;
            const double a = fma(c.kf3, g15, 1.);
            const double n2 = 0.0065 * c.kf1 * g02;
            const double b = fma(c.kf2, g20, 1.);
            const double r = frcp3(a * b);
            f = gatef(f, dt, n1 * b * r, fma(n1, b, n2 * a) * r);
        }
        io.st(3, d); io.st(4, f);
        io.st(6, fma(dt, fma(-0.0001, I_Si, 0.07 * (0.0001 - cai)), cai));
        // ---- I_K (calc_ik :297-356): old x
        const double Xi = u > -100. ? 2.837 * fma(-c.kxb, g04, c.kxa) * frcp3(u + 77.) : 1.;
        double x = io.ld(5);
        const double I_K = c.G_K * x * Xi * (u - c.E_K);
        {
            const double n1 = 0.0005 * EX(0.083 * (u + 50.));
            const double a = 1. + EX(0.057 * (u + 50.));
            const double n2 = 0.0013 * c.kx1 * g06;
            const double b = fma(c.kx2, g04, 1.);
            const double r = frcp3(a * b);
            x = gatef(x, dt, n1 * b * r, fma(n1, b, n2 * a) * r);
        }
        io.st(5, x);
        // ---- I_K1, I_Kp, I_b (calc_ik1 :359-408, calc_ikp :411-427, calc_ib :430-443)
        const double y = u - c.E_K1;
        const double a1 = 1. + EXC(0.2385 * (y - 59.215));
        const double n1 = fma(0.49124, EXC(0.08032 * (y + 5.476)),
                              EXC(0.06175 * (y - 594.31)));
        const double b1 = 1.02 * (1. + EXC(-0.5143 * (y + 4.753)));
        const double ik1 = c.G_K1 * (b1 * frcp3(fma(n1, a1, b1))) * y;
        const double ikp = c.gkp * frcp3(1. + EX((7.488 - u) * (1. / 5.98))) * y;
        const double ib = c.gb * (u + 59.87);
        un -= dt * (ina + I_Si + (ik1 + ikp + ib) + I_K);
    }

    template <class IO, class E>
    FWB_HD static void ionic_impl(double u, double &un, IO &io, const Consts &c, double cai)
    {
        const double dt = c.dt;
        // calc_ina :185-241
        double alpha_h = 0, beta_h = 0, beta_J = 0, alpha_J = 0;
        if (u >= -40.) {
            beta_h = 1. / (0.13 * (1 + E::eu(FWB_EDIVK(u + 10.66, -11.1))));
            beta_J = E::dv(0.3 * E::eu(-2.535 * 1e-07 * u), 1 + E::eu(-0.1 * (u + 32)));
        } else {
            alpha_h = 0.135 * E::eu(FWB_EDIVK(80 + u, -6.8));
            beta_h = 3.56 * E::eu(0.079 * u) + 3.1 * 1e5 * E::eu(0.35 * u);
            beta_J = E::dv(0.1212 * E::eu(-0.01052 * u), 1 + E::eu(-0.1378 * (u + 40.14)));
            alpha_J = E::dv((-1.2714 * 1e5 * E::eu(0.2444 * u) - 3.474 * 1e-5 * E::eu(-0.04391 * u)) *
                                (u + 37.78),
                            1 + E::eu(0.311 * (u + 79.23)));
        }
        const double alpha_m = 0.32 * (u + 47.13) / (1 - E::eu(-0.1 * (u + 47.13)));
        const double beta_m = 0.08 * E::eu(FWB_EDIVK(-u, 11.));
        const double m = gate(io.ld(0), dt, alpha_m, beta_m);
        const double h = gate(io.ld(1), dt, alpha_h, beta_h);
        const double j = gate(io.ld(2), dt, alpha_J, beta_J);
        io.st(0, m); io.st(1, h); io.st(2, j);
        const double ina = c.gna * m * m * m * h * j * (u - c.E_Na);
        // calc_isk :244-294
        double d = io.ld(3), f = io.ld(4);
        const double E_Si = 7.7 - 13.0287 * log(cai);
        const double I_Si = c.gsi * d * f * (u - E_Si);
        const double alpha_d = E::dv(0.095 * E::eu(-0.01 * (u - 5)), 1 + E::eu(-0.072 * (u - 5)));
        const double beta_d = E::dv(0.07 * E::eu(-0.017 * (u + 44)), 1 + E::eu(0.05 * (u + 44)));
        const double alpha_f = E::dv(0.012 * E::eu(-0.008 * (u + 28)), 1 + E::eu(0.15 * (u + 28)));
        const double beta_f = E::dv(0.0065 * E::eu(-0.02 * (u + 30)), 1 + E::eu(-0.2 * (u + 30)));
        d = gate(d, dt, alpha_d, beta_d);
        f = gate(f, dt, alpha_f, beta_f);
        cai += dt * (-0.0001 * I_Si + 0.07 * (0.0001 - cai));
        io.st(3, d); io.st(4, f); io.st(6, cai);
        // calc_ik :297-356
        double Xi;
        if (u > -100)
            Xi = 2.837 * (E::eu(0.04 * (u + 77)) - 1) / ((u + 77) * E::eu(0.04 * (u + 35)));
        else
            Xi = 1;
        double x = io.ld(5);
        const double I_K = c.G_K * x * Xi * (u - c.E_K);
        const double alpha_x = E::dv(0.0005 * E::eu(0.083 * (u + 50)), 1 + E::eu(0.057 * (u + 50)));
        const double beta_x = E::dv(0.0013 * E::eu(-0.06 * (u + 20)), 1 + E::eu(-0.04 * (u + 20)));
        x = gate(x, dt, alpha_x, beta_x);
        io.st(5, x);
        // calc_ik1 :359-408, calc_ikp :411-427, calc_ib :430-443, kernel :496-510
        const double E_K1 = c.E_K1;
        const double alpha_K1 = E::dv(1.02, 1 + E::eu(0.2385 * (u - E_K1 - 59.215)));
        const double beta_K1 = (0.49124 * E::eu(0.08032 * (u - E_K1 + 5.476)) +
                                E::eu(0.06175 * (u - E_K1 - 594.31))) /
                               (1 + E::eu(-0.5143 * (u - E_K1 + 4.753)));
        const double K_1x = alpha_K1 / (alpha_K1 + beta_K1);
        const double ik1 = c.G_K1 * K_1x * (u - E_K1);
        const double K_p = E::dv(1., 1 + E::eu(FWB_EDIVK(7.488 - u, 5.98)));
        const double ikp = c.gkp * K_p * (u - E_K1);
        const double ib = c.gb * (u + 59.87);
        const double ik1t = ik1 + ikp + ib;
        un -= dt * (ina + I_Si + ik1t + I_K);
    }
};

// ---------------------------------------------------------------------------
// ten Tusscher-Panfilov 2006 -- cpuwave2D/model/tp06_2d.py:242-984 (point
// functions), :987-1103 (ionic_kernel_2d); cpuwave3D/model/tp06_3d.py:85-204.
// Voltage gates + fcass: Rush-Larsen; rr: forward Euler; oo algebraic; casr,
// cass: analytic buffer quadratic; cai: NOT updated (tp06_2d.py:1098 assigns
// the old value last); oo's previous value is never read.
// state: 0 cai, 1 casr, 2 cass, 3 nai, 4 Ki, 5 m, 6 h, 7 j, 8 xr1, 9 xr2,
//        10 xs, 11 r, 12 s, 13 d, 14 f, 15 f2, 16 fcass, 17 rr, 18 oo
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_TP06> {
#ifndef FWB_TP06_MIN_BLOCKS
#define FWB_TP06_MIN_BLOCKS 3      // 80 registers: no spills in any instantiation (4 -> 64 regs spills 90-150 B)
#endif
    static constexpr int NS = 19, NP = 49, MIN_BLOCKS = FWB_TP06_MIN_BLOCKS;
    static constexpr bool USE_TMA = false;   // step_kernel_tma for HBM-bound models
    static constexpr uint32_t READ_MASK = 0x3ffff;          // all but oo
    static constexpr uint32_t WRITE_MASK = 0x7ffff & ~0x1u;  // all but cai
    struct Consts {
        double dt, ko, cao, nao, RTONF, half_RTONF, CAPACITANCE;
        double pKNa_nao, ko_pKNa_nao, pKNa;
        double gna, gcal, gto, gkr_sqrt, gks, gk1, gpca, KpCa, gpk, gbna, gbca;
        double F, FF_RT;
        DivC RT;
        double inaca_pref, ksat, n_, n_m1, knak_pref, KmNa;
        double Bufsr, Kbufsr, Bufss, Kbufss, Vmaxup, Kup2, Vrel, k1_, k2_, k3, k4, EC;
        double maxsr, maxsr_m_minsr, Vleak, Vxfer, Vc_Vss, Vsr_Vss;
        double inverseVcF, inversevssF2;
        // fast path only (ionic_fast): logs of the numerators of the reversal potentials,
        // EC^2, F / (R T), and exp(offset / slope) of every shared-slope exponential
        double l_ko, l_nao, l_kpn, l_cao, EC2, F_RT;
        int fast_ok;       // parameters allow the rearranged path (positive concentrations)
        double km1, km2, kj1, kd1, kd2, kd3, kf1, kf2, kf3, kf4, kf5, kr1, ks1, ks2;
        double kx1, kx2, kx3, kx4, kx5, kxs1, kxs2, kxs3;
        double ec[8];      // fexp's reduction / polynomial constants (fexp_fill_consts)
    };
    static bool derive(const double *p, double dt, Consts &c)
    {
        derive_fast(p, c);
        fexp_fill_consts(c.ec);
        const double ko = p[0], cao = p[1], nao = p[2], Vc = p[3], Vsr = p[4], Vss = p[5],
                     R = p[24], F = p[25], T = p[26], gkr = p[29], pKNa = p[30],
                     KmK = p[34], knak = p[36], knaca = p[39], KmNai = p[40], KmCa = p[41];
        c.dt = dt; c.ko = ko; c.cao = cao; c.nao = nao;
        c.RTONF = p[27]; c.half_RTONF = 0.5 * p[27]; c.CAPACITANCE = p[28];
        c.pKNa = pKNa; c.pKNa_nao = pKNa * nao; c.ko_pKNa_nao = ko + pKNa * nao;
        c.gna = p[32]; c.gcal = p[37]; c.gto = p[47];
        c.gkr_sqrt = gkr * sqrt(ko / 5.4);
        c.gks = p[48]; c.gk1 = p[31]; c.gpca = p[44]; c.KpCa = p[45]; c.gpk = p[46];
        c.gbna = p[33]; c.gbca = p[38];
        c.F = F; c.RT = make_divc(R * T); c.FF_RT = F * F / (R * T);
        c.inaca_pref = knaca * (1. / (KmNai * KmNai * KmNai + nao * nao * nao)) * (1. / (KmCa + cao));
        c.ksat = p[42]; c.n_ = p[43]; c.n_m1 = p[43] - 1;
        c.knak_pref = knak * (ko / (ko + KmK)); c.KmNa = p[35];
        c.Bufsr = p[8]; c.Kbufsr = p[9]; c.Bufss = p[10]; c.Kbufss = p[11];
        c.Vmaxup = p[12]; c.Kup2 = p[13] * p[13]; c.Vrel = p[14]; c.k1_ = p[15];
        c.k2_ = p[16]; c.k3 = p[17]; c.k4 = p[18]; c.EC = p[19];
        c.maxsr = p[20]; c.maxsr_m_minsr = p[20] - p[21]; c.Vleak = p[22]; c.Vxfer = p[23];
        c.Vc_Vss = Vc / Vss; c.Vsr_Vss = Vsr / Vss;
        c.inverseVcF = 1. / (Vc * F);
        c.inversevssF2 = 1. / (2 * Vss * F);
        return divc_ok(c.RT);
    }
    static void derive_fast(const double *p, Consts &c)
    {
        const double ko = p[0], cao = p[1], nao = p[2], R = p[24], F = p[25], T = p[26],
                     pKNa = p[30];
        c.l_ko = log(ko); c.l_nao = log(nao); c.l_kpn = log(ko + pKNa * nao); c.l_cao = log(cao);
        c.EC2 = p[19] * p[19]; c.F_RT = F / (R * T);
        c.fast_ok = ko > 0 && nao > 0 && cao > 0 && pKNa >= 0 && ko < 1e300 && nao < 1e300 &&
                    cao < 1e300 && pKNa < 1e300;
        c.km1 = exp(-60. / 5.); c.km2 = exp(35. / 5.); c.kj1 = exp(-0.1 * 32.);
        c.kd1 = exp(-8. / 7.5); c.kd2 = exp(5. / 5.); c.kd3 = exp(50. / 20.);
        c.kf1 = exp(20. / 7.); c.kf2 = exp(13. / 10.); c.kf3 = exp(30. / 10.);
        c.kf4 = exp(35. / 7.); c.kf5 = exp(25. / 10.);
        c.kr1 = exp(20. / 6.); c.ks1 = exp(20. / 5.); c.ks2 = exp(-20. / 5.);
        c.kx1 = exp(-26. / 7.); c.kx2 = exp(-45. / 10.); c.kx3 = exp(88. / 24.);
        c.kx4 = exp(-60. / 20.); c.kx5 = exp(-60. / 20.);
        c.kxs1 = exp(-5. / 14.); c.kxs2 = exp(5. / 6.); c.kxs3 = exp(-35. / 15.);
    }
    // Rush-Larsen update (tp06_2d.py:307-309 and twins)
    template <class E> FWB_HD static double rl(double inf, double x, double dt, double tau)
    {
        return inf - (inf - x) * E::en(-dt / tau);
    }
    // the same for tau = 1 / rate (h and j gates, tp06_2d.py:296-309): the reference
    // divides twice, -dt / (1.0 / rate); FastMath multiplies once
    template <class E> FWB_HD static double rl_rate(double inf, double x, double dt, double rate)
    {
        if (E::FUSE_RATE) return inf - (inf - x) * E::en(-dt * rate);
        return inf - (inf - x) * E::en(-dt / (1.0 / rate));
    }
    // The rearranged path (ionic_fast) needs |u| < 300 mV (fexp) and positive, normal
    // concentrations (flog, branch-free reciprocals); anything else -- NaNs included, they
    // fail every comparison -- takes the reference statement (ionic_ref).  The tile kernel
    // asks fast_ok() for the whole block and hands tiles with an ineligible node to a second,
    // cold kernel, so the hot kernel carries the fast path only.
    static constexpr bool HAS_FAST = true;
    template <class IO> FWB_HD static bool fast_ok(double u, const IO &io, const Consts &c)
    {
        return c.fast_ok && fabs(u) < FAST_MATH_U_LIMIT && conc_ok(io.ld(0)) &&
               conc_ok(io.ld(3)) && conc_ok(io.ld(4));
    }
    template <class IO>
    FWB_HD static void ionic_fastpath(double u, double &un, IO &io, const Consts &c)
    {
        ionic_fast(u, un, io, c, io.ld(0), io.ld(3), io.ld(4));
    }
    template <class IO>
    FWB_HD static void ionic_ref(double u, double &un, IO &io, const Consts &c)
    {
        ionic_impl<IO, LibMath>(u, un, io, c, io.ld(0), io.ld(3), io.ld(4));
    }
    template <class IO>
    FWB_HD static void ionic(double u, double &un, IO &io, const Consts &c)
    {
#ifdef __CUDA_ARCH__
        if (fast_ok(u, io, c)) ionic_fastpath(u, un, io, c);
        else
#endif
            ionic_ref(u, un, io, c);
    }
    FWB_HD static bool conc_ok(double x) { return x > 1e-300 && x < 1e300; }

    // ------------------------------------------------------------------------------------
    // The device's normal path for |u| < 300 mV: the same equations as ionic_impl (which is
    // the reference statement, line by line), rearranged for the FP64 pipe.  Every change is
    // an algebraic identity, so the two differ by rounding only (tests/test_host_models.py
    // compares them on random node states; the GPU parity tests hold the 1e-9 bar):
    //  * exponentials of u with commensurable slopes share one evaluation:
    //    exp((u + a) / s) = exp(a / s) * exp(u / s), exp(u / 10) = exp(u / 20)^2, ...;
    //    1 / (1 + K / g) is evaluated as g / (g + K)
    //  * a Rush-Larsen factor needs 1 / tau, never tau: the sums and products of
    //    k / (1 + exp) terms that make up tau are put over one denominator, so each gate
    //    costs one reciprocal instead of one per term plus a division
    //  * every remaining division has a denominator that is positive and normal for any
    //    physical state: reciprocal seed + one third-order Newton step (frcp3, <= 1 ulp)
    //    without the special-operand branch
    //  * log(a / x) = log(a) - log(x) with log(a) from the host and a table-driven flog(x)
    //    (fexp.cuh: 10 FP64-pipe instructions instead of ~30 + 60 others)
    // ------------------------------------------------------------------------------------
    FWB_HD static double rlf(double inf, double x, double dt, double rtau)
    {
        return fma(x - inf, fexp_fast_neg(-dt * rtau), inf);
    }
    template <class IO>
    FWB_HD static void ionic_fast(double u, double &un, IO &io, const Consts &c, double cai,
                                  double nai, double Ki)
    {
        const double dt = c.dt;
        const double *ec = c.ec;       // kernel-parameter bank -> uniform registers
        auto EX = [ec](double x) { return fexp_fast_p(x, ec); };
        auto EXC = [ec](double x) {
            x = x < -700.0 ? -700.0 : x;
            return fexp_fast_p(x > 700.0 ? 700.0 : x, ec);
        };
        auto RL = [ec](double inf, double x, double dt_, double rtau) {
            const double a = -dt_ * rtau;
            return fma(x - inf, fexp_fast_p(a < -700.0 ? -700.0 : a, ec), inf);
        };
        const double Ek = c.RTONF * (c.l_ko - flog(Ki));
        const double Ena = c.RTONF * (c.l_nao - flog(nai));
        const double Eks = c.RTONF * (c.l_kpn - flog(fma(c.pKNa, nai, Ki)));
        const double Eca = c.half_RTONF * (c.l_cao - flog(cai));

        // shared exponentials of u
        const double g20 = EX(u * 0.05), g10 = g20 * g20, g5 = g10 * g10;
        const double i20 = frcp3(g20), i10 = i20 * i20;
        const double g14 = EX(u * (1. / 14.)), g7 = g14 * g14;
        const double g15 = EX(u * (1. / 15.)), g75 = g15 * g15;
        const double n24 = EX(u * (-1. / 24.)), n12 = n24 * n24, n6 = n12 * n12;

        // the twelve membrane currents are accumulated as they appear (sum for u_new, the
        // sodium and the potassium balance) instead of being kept live to the end
        double s_u, s_na, s_k;
        // ---- I_Na (calc_ina :242-318)
        {
            const double A = g5 + c.km1;                  // alpha_m = g5 / A
            const double B = fma(c.km2, g5, 1.);          // beta_m = 0.1 / B + 0.1 / C
            const double C = 1. + EX((u - 50.) * (1. / 200.));
            const double rtau_m = A * B * C * frcp3(0.1 * g5 * (B + C));
            const double em = 1. + EX((-56.86 - u) * (1. / 9.03));
            const double m_inf = frcp3(em * em);
            const double eh = 1. + EX((u + 71.55) * (1. / 7.43));
            const double h_inf = frcp3(eh * eh);
            double rate_h, rate_j;
            if (u >= -40.) {
                rate_h = 0.77 * frcp3(0.13 * (1. + EX((u + 10.66) * (-1. / 11.1))));
                rate_j = 0.6 * EX(0.057 * u) * g10 * frcp3(g10 + c.kj1);
            } else {
                rate_h = 0.057 * EX((u + 80.) * (-1. / 6.8)) +
                         (2.7 * EX(0.079 * u) + 3.1e5 * EX(0.3485 * u));
                const double a = 1. + EX(0.311 * (u + 79.23));
                const double b = 1. + EX(-0.1378 * (u + 40.14));
                const double na = (-2.5428e4 * EX(0.2444 * u) -
                                   6.948e-6 * EX(-0.04391 * u)) * (u + 37.78);
                const double nb = 0.02424 * EX(-0.01052 * u);
                rate_j = fma(na, b, nb * a) * frcp3(a * b);
            }
            const double m = RL(m_inf, io.ld(5), dt, rtau_m);
            const double h = RL(h_inf, io.ld(6), dt, rate_h);
            const double j = RL(h_inf, io.ld(7), dt, rate_j);
            io.st(5, m); io.st(6, h); io.st(7, j);
            const double uEna = u - Ena;
            const double ina = c.gna * m * m * m * h * j * uEna;
            const double ibna = c.gbna * uEna;
            s_na = ina + ibna;
            s_u = s_na + c.gbca * (u - Eca);          // + I_bCa (:673-697)
        }

        // ---- I_CaL (calc_ical :321-380)
        const double cass = io.ld(2);
        double ical;
        {
            const double d_inf = g75 * frcp3(g75 + c.kd1);
            const double A = 1. + EX((-35. - u) * (1. / 13.));   // Ad = 1.4 / A + 0.25
            const double B = fma(c.kd2, g5, 1.);                        // Bd = 1.4 / B
            const double Cn = g20 + c.kd3;                              // Cd = g20 / Cn
            const double P = fma(0.25, A, 1.4) * 1.4;                   // Ad Bd = P / (A B)
            const double AB = A * B;
            const double rtau_d = AB * Cn * frcp3(fma(P, Cn, g20 * AB));
            const double d = RL(d_inf, io.ld(13), dt, rtau_d);
            io.st(13, d);

            const double E30 = fma(c.kf3, g10, 1.);
            const double f_inf = frcp3(fma(c.kf1, g7, 1.));
            const double Af = 1102.5 * EX(-(u + 27.) * (u + 27.) * (1. / 225.));
            const double Bn = g10 + c.kf2;                              // Bf = 200 g10 / Bn
            const double Df = Bn * E30;                                 // Cf = 180 / E30 + 20
            const double Nf = fma(200. * g10, E30, 180. * Bn);
            const double rtau_f = Df * frcp3(fma(Af + 20., Df, Nf));
            const double f = RL(f_inf, io.ld(14), dt, rtau_f);
            io.st(14, f);

            const double f2_inf = fma(0.67, frcp3(fma(c.kf4, g7, 1.)), 0.33);
            const double Af2 = 600. * EX(-(u + 25.) * (u + 25.) * (1. / 170.));
            const double Bn2 = g10 + c.kf5;                             // Bf2 = 31 g10 / Bn2
            const double Df2 = Bn2 * E30;                               // Cf2 = 16 / E30
            const double Nf2 = fma(31. * g10, E30, 16. * Bn2);
            const double rtau_f2 = Df2 * frcp3(fma(Af2, Df2, Nf2));
            const double f2 = RL(f2_inf, io.ld(15), dt, rtau_f2);
            io.st(15, f2);

            const double cs = cass * (1. / 0.05);
            const double cq = fma(cs, cs, 1.);
            const double fcass_inf = fma(0.6, frcp3(cq), 0.4);
            const double rtau_fcass = cq * frcp3(fma(2., cq, 80.));
            const double fcass = RL(fcass_inf, io.ld(16), dt, rtau_fcass);
            io.st(16, fcass);

            const double e2 = EX(2. * (u - 15.) * c.F_RT);
            ical = c.gcal * d * f * f2 * fcass * 4. * (u - 15.) * c.FF_RT *
                   fma(0.25 * e2, cass, -c.cao) * frcp3(e2 - 1.);
            s_u += ical;
        }

        // ---- I_to (calc_ito :383-413)
        const double y = u - Ek;
        {
            const double r_inf = frcp3(fma(c.kr1, n6, 1.));
            const double s_inf = frcp3(fma(c.ks1, g5, 1.));
            const double rtau_r =
                frcp3(fma(9.5, EX(-(u + 40.) * (u + 40.) * (1. / 1800.)), 0.8));
            const double G = fma(85., EX(-(u + 45.) * (u + 45.) * (1. / 320.)), 3.);
            const double S2 = fma(c.ks2, g5, 1.);                       // tau_s = G + 5 / S2
            const double rtau_s = S2 * frcp3(fma(G, S2, 5.));
            const double sg = RL(s_inf, io.ld(12), dt, rtau_s);
            const double r = RL(r_inf, io.ld(11), dt, rtau_r);
            io.st(11, r); io.st(12, sg);
            s_k = c.gto * r * sg * y;
        }

        // ---- I_Kr (calc_ikr :416-452)
        {
            const double xr1_inf = g7 * frcp3(g7 + c.kx1);
            // tau_xr1 = 450 g10 / (g10 + kx2) * 6 / (1 + e115)
            const double e115 = EX((u + 30.) * (1. / 11.5));
            const double rtau_xr1 = (g10 + c.kx2) * (1. + e115) * (i10 * (1. / 2700.));
            const double xr2_inf = n24 * frcp3(n24 + c.kx3);
            // tau_xr2 = 3 g20 / (g20 + kx4) * 1.12 / (1 + kx5 g20)
            const double rtau_xr2 = (g20 + c.kx4) * fma(c.kx5, g20, 1.) * (i20 * (1. / 3.36));
            const double xr1 = RL(xr1_inf, io.ld(8), dt, rtau_xr1);
            const double xr2 = RL(xr2_inf, io.ld(9), dt, rtau_xr2);
            io.st(8, xr1); io.st(9, xr2);
            s_k += c.gkr_sqrt * xr1 * xr2 * y;
        }

        // ---- I_Ks (calc_iks :455-485)
        {
            const double xs_inf = g14 * frcp3(g14 + c.kxs1);
            const double sq = fsqrt(fma(c.kxs2, n6, 1.));                // Axs = 1400 / sq
            const double sb = sq * fma(c.kxs3, g15, 1.);                // Bxs = 1 / (1 + kxs3 g15)
            const double rtau_xs = sb * frcp3(fma(80., sb, 1400.));      // tau_xs = 1400 / sb + 80
            const double xs = RL(xs_inf, io.ld(10), dt, rtau_xs);
            io.st(10, xs);
            s_k += c.gks * xs * xs * (u - Eks);
        }

        // ---- I_K1 (calc_ik1 :488-514): rec = ak1 / (ak1 + bk1), ak1 = 0.1 / a, bk1 = n / b
        {
            const double a = 1. + EXC(0.06 * (y - 200.));
            const double b = 0.1 * (1. + EXC(-0.5 * y));
            const double n = fma(3., EXC(0.0002 * (y + 100.)),
                                 EXC(0.1 * (y - 10.)));
            s_k += c.gk1 * (b * frcp3(fma(n, a, b))) * y;
        }
        // ---- I_NaCa (calc_inaca :517-565), I_NaK (calc_inak :568-604)
        const double x = u * c.F_RT;
        const double e_nm1 = EX(c.n_m1 * x);
        const double en01 = EX(-0.1 * x);
        const double en02 = en01 * en01, en04 = en02 * en02, en08 = en04 * en04;
        const double en_x = en08 * en02;                                // exp(-x)
        const double nai3 = nai * nai * nai;
        const double inaca = c.inaca_pref * e_nm1 *
                             fma(-2.5 * (c.nao * c.nao * c.nao) * cai, en_x, nai3 * c.cao) *
                             frcp3(en_x * fma(c.ksat, e_nm1, 1.));
        const double inak = c.knak_pref * nai *
                            frcp3((nai + c.KmNa) * fma(0.0353, en_x, fma(0.1245, en01, 1.)));
        // ---- I_pCa, I_pK (:607-670)
        const double ipca = c.gpca * cai * frcp3(c.KpCa + cai);
        const double ipk = c.gpk * frcp3(1. + EX((25. - u) * (1. / 5.98))) * y;

        // calc_nai :928-953, calc_ki :956-984 (old nai, Ki)
        {
            const double both = inak + inaca;
            s_u += s_k + ipk + both + ipca;
            const double dNai = -fma(3., both, s_na) * c.inverseVcF * c.CAPACITANCE;
            io.st(3, fma(dt, dNai, nai));
            const double dKi = -(s_k + fma(-2., inak, ipk)) * c.inverseVcF * c.CAPACITANCE;
            io.st(4, fma(dt, dKi, Ki));
        }
        un -= dt * s_u;

        // ---- calcium handling (calc_irel :700-730 ... calc_cass :831-872)
        const double casr = io.ld(1);
        const double cass2 = cass * cass;
        double irel;
        {
            double rr = io.ld(17);
            const double casr2 = casr * casr;
            const double kCaSR = fma(-c.maxsr_m_minsr * casr2, frcp3(casr2 + c.EC2), c.maxsr);
            const double k2 = c.k2_ * kCaSR;
            rr = fma(dt, fma(c.k4, 1. - rr, -(k2 * cass * rr)), rr);
            const double k1c = c.k1_ * cass2;
            const double oo = k1c * rr * frcp3(fma(c.k3, kCaSR, k1c));
            irel = c.Vrel * oo * (casr - cass);
            io.st(17, rr); io.st(18, oo);
        }
        const double ileak = c.Vleak * (casr - cai);
        const double cai2 = cai * cai;
        const double iup = c.Vmaxup * cai2 * frcp3(cai2 + c.Kup2);
        const double ixfer = c.Vxfer * (cass - cai);
        {
            const double CaCSQN = c.Bufsr * casr * frcp3(casr + c.Kbufsr);
            const double dCaSR = dt * (iup - irel - ileak);
            const double bjsr = c.Bufsr - CaCSQN - dCaSR - casr + c.Kbufsr;
            const double cjsr = c.Kbufsr * (CaCSQN + dCaSR + casr);
            io.st(1, (fsqrt(fma(bjsr, bjsr, 4 * cjsr)) - bjsr) * 0.5);
        }
        {
            const double CaSSBuf = c.Bufss * cass * frcp3(cass + c.Kbufss);
            const double dCaSS = dt * (-ixfer * c.Vc_Vss + irel * c.Vsr_Vss +
                                       (-ical * c.inversevssF2 * c.CAPACITANCE));
            const double bcss = c.Bufss - CaSSBuf - dCaSS - cass + c.Kbufss;
            const double ccss = c.Kbufss * (CaSSBuf + dCaSS + cass);
            io.st(2, (fsqrt(fma(bcss, bcss, 4 * ccss)) - bcss) * 0.5);
        }
    }

    template <class IO, class E>
    FWB_HD static void ionic_impl(double u, double &un, IO &io, const Consts &c, double cai,
                                  double nai, double Ki)
    {
        const double dt = c.dt;
        // reversal potentials :1072-1075
        const double Ek = c.RTONF * log(c.ko / Ki);
        const double Ena = c.RTONF * log(c.nao / nai);
        const double Eks = c.RTONF * log(c.ko_pKNa_nao / (Ki + c.pKNa * nai));
        const double Eca = c.half_RTONF * log(c.cao / cai);

        // running sums in the reference's final association orders:
        //   itot = ikr+iks+ik1+ito+ina+ibna+ical+ibca+inak+inaca+ipca+ipk   (:1103)
        //   dNai: ina+ibna+3 inak+3 inaca        dKi: ik1+ito+ikr+iks-2 inak+ipk
        // calc_ina :242-318
        double ina;
        {
            const double alpha_m = E::dv(1., 1. + E::eu(FWB_EDIVK(-60. - u, 5.)));
            const double beta_m = E::dv(0.1, 1. + E::eu(FWB_EDIVK(u + 35., 5.))) +
                                  E::dv(0.10, 1. + E::eu(FWB_EDIVK(u - 50., 200.)));
            const double tau_m = alpha_m * beta_m;
            const double em = 1. + E::eu(FWB_EDIVK(-56.86 - u, 9.03));
            const double m_inf = E::dv(1., em * em);
            double alpha_h, beta_h, alpha_j, beta_j;
            if (u >= -40.) {
                alpha_h = 0.;
                beta_h = E::dv(0.77, 0.13 * (1. + E::eu(FWB_EDIVK(-(u + 10.66), 11.1))));
                alpha_j = 0.;
                beta_j = E::dv(0.6 * E::eu(0.057 * u), 1. + E::eu(-0.1 * (u + 32.)));
            } else {
                alpha_h = 0.057 * E::eu(FWB_EDIVK(-(u + 80.), 6.8));
                beta_h = 2.7 * E::eu(0.079 * u) + 3.1e5 * E::eu(0.3485 * u);
                alpha_j = E::dv((-2.5428e4 * E::eu(0.2444 * u) - 6.948e-6 * E::eu(-0.04391 * u)) *
                                    (u + 37.78),
                                1. + E::eu(0.311 * (u + 79.23)));
                beta_j = E::dv(0.02424 * E::eu(-0.01052 * u), 1. + E::eu(-0.1378 * (u + 40.14)));
            }
            const double eh = 1. + E::eu(FWB_EDIVK(u + 71.55, 7.43));
            const double h_inf = E::dv(1., eh * eh);
            const double m = rl<E>(m_inf, io.ld(5), dt, tau_m);
            const double h = rl_rate<E>(h_inf, io.ld(6), dt, alpha_h + beta_h);   // tau_h = 1/(..)
            const double j = rl_rate<E>(h_inf, io.ld(7), dt, alpha_j + beta_j);   // tau_j = 1/(..)
            io.st(5, m); io.st(6, h); io.st(7, j);
            ina = c.gna * m * m * m * h * j * (u - Ena);
        }

        // calc_ical :321-380
        const double cass = io.ld(2);
        double ical;
        {
            const double d_inf = E::dv(1., 1. + E::eu(FWB_EDIVK(-8 - u, 7.5)));
            const double Ad = E::dv(1.4, 1. + E::eu(FWB_EDIVK(-35 - u, 13.))) + 0.25;
            const double Bd = E::dv(1.4, 1. + E::eu(FWB_EDIVK(u + 5, 5.)));
            const double Cd = E::dv(1., 1. + E::eu(FWB_EDIVK(50 - u, 20.)));
            const double tau_d = Ad * Bd + Cd;
            const double d = rl<E>(d_inf, io.ld(13), dt, tau_d);
            io.st(13, d);
            const double f_inf = E::dv(1., 1. + E::eu(FWB_EDIVK(u + 20, 7.)));
            const double Af = 1102.5 * E::eu(FWB_EDIVK(-(u + 27) * (u + 27), 225.));
            const double Bf = E::dv(200., 1 + E::eu(FWB_EDIVK(13 - u, 10.)));
            const double e30 = E::eu(FWB_EDIVK(u + 30, 10.));
            const double Cf = E::dv(180., 1 + e30) + 20;
            const double tau_f = Af + Bf + Cf;
            const double f = rl<E>(f_inf, io.ld(14), dt, tau_f);
            io.st(14, f);
            const double f2_inf = E::dv(0.67, 1. + E::eu(FWB_EDIVK(u + 35, 7.))) + 0.33;
            const double Af2 = 600 * E::eu(FWB_EDIVK(-(u + 25) * (u + 25), 170.));
            const double Bf2 = E::dv(31., 1. + E::eu(FWB_EDIVK(25 - u, 10.)));
            const double Cf2 = E::dv(16., 1. + e30);
            const double tau_f2 = Af2 + Bf2 + Cf2;
            const double f2 = rl<E>(f2_inf, io.ld(15), dt, tau_f2);
            io.st(15, f2);
            const double cq = 1 + FWB_EDIVK(cass, 0.05) * FWB_EDIVK(cass, 0.05);
            const double fcass_inf = E::dv(0.6, cq) + 0.4;
            const double tau_fcass = E::dv(80., cq) + 2.;
            const double fcass = rl<E>(fcass_inf, io.ld(16), dt, tau_fcass);
            io.st(16, fcass);
            const double e2 = E::eu(E::dk(2 * (u - 15) * c.F, c.RT));
            ical = c.gcal * d * f * f2 * fcass * 4 * (u - 15) * c.FF_RT *
                   (0.25 * e2 * cass - c.cao) / (e2 - 1.);
        }

        // calc_ito :383-413
        double ito;
        {
            const double r_inf = E::dv(1., 1. + E::eu(FWB_EDIVK(20 - u, 6.)));
            const double s_inf = E::dv(1., 1. + E::eu(FWB_EDIVK(u + 20, 5.)));
            const double tau_r = 9.5 * E::eu(FWB_EDIVK(-(u + 40.) * (u + 40.), 1800.)) + 0.8;
            const double tau_s = 85. * E::eu(FWB_EDIVK(-(u + 45.) * (u + 45.), 320.)) +
                                 E::dv(5., 1. + E::eu(FWB_EDIVK(u - 20., 5.))) + 3.;
            const double sg = rl<E>(s_inf, io.ld(12), dt, tau_s);
            const double r = rl<E>(r_inf, io.ld(11), dt, tau_r);
            io.st(11, r); io.st(12, sg);
            ito = c.gto * r * sg * (u - Ek);
        }

        // calc_ikr :416-452
        double ikr;
        {
            const double xr1_inf = E::dv(1., 1. + E::eu(FWB_EDIVK(-26. - u, 7.)));
            const double axr1 = E::dv(450., 1. + E::eu(FWB_EDIVK(-45. - u, 10.)));
            const double bxr1 = E::dv(6., 1. + E::eu(FWB_EDIVK(u - (-30.), 11.5)));
            const double tau_xr1 = axr1 * bxr1;
            const double xr2_inf = E::dv(1., 1. + E::eu(FWB_EDIVK(u - (-88.), 24.)));
            const double axr2 = E::dv(3., 1. + E::eu(FWB_EDIVK(-60. - u, 20.)));
            const double bxr2 = E::dv(1.12, 1. + E::eu(FWB_EDIVK(u - 60., 20.)));
            const double tau_xr2 = axr2 * bxr2;
            const double xr1 = rl<E>(xr1_inf, io.ld(8), dt, tau_xr1);
            const double xr2 = rl<E>(xr2_inf, io.ld(9), dt, tau_xr2);
            io.st(8, xr1); io.st(9, xr2);
            ikr = c.gkr_sqrt * xr1 * xr2 * (u - Ek);
        }

        // calc_iks :455-485
        double iks;
        {
            const double xs_inf = E::dv(1., 1. + E::eu(FWB_EDIVK(-5. - u, 14.)));
            const double Axs = (1400. / (sqrt(1. + E::eu(FWB_EDIVK(5. - u, 6.)))));
            const double Bxs = E::dv(1., 1. + E::eu(FWB_EDIVK(u - 35., 15.)));
            const double tau_xs = Axs * Bxs + 80;
            const double xs = rl<E>(xs_inf, io.ld(10), dt, tau_xs);
            io.st(10, xs);
            iks = c.gks * xs * xs * (u - Eks);
        }

        // calc_ik1 :488-514
        const double ak1 = E::dv(0.1, 1. + E::ec(0.06 * (u - Ek - 200)));
        const double bk1 = E::dv(3. * E::ec(0.0002 * (u - Ek + 100)) + E::ec(0.1 * (u - Ek - 10)),
                                 1. + E::ec(-0.5 * (u - Ek)));
        const double rec_iK1 = ak1 / (ak1 + bk1);
        const double ik1 = c.gk1 * rec_iK1 * (u - Ek);
        // calc_inaca :517-565
        const double e_nm1 = E::eu(E::dk(c.n_m1 * u * c.F, c.RT));
        const double inaca = c.inaca_pref * (1. / (1 + c.ksat * e_nm1)) *
                             (E::eu(E::dk(c.n_ * u * c.F, c.RT)) * nai * nai * nai * c.cao -
                              e_nm1 * c.nao * c.nao * c.nao * cai * 2.5);
        // calc_inak :568-604
        const double rec_iNaK = E::dv(1., 1. + 0.1245 * E::eu(E::dk(-0.1 * u * c.F, c.RT)) +
                                              0.0353 * E::eu(E::dk(-u * c.F, c.RT)));
        const double inak = c.knak_pref * (nai / (nai + c.KmNa)) * rec_iNaK;
        // calc_ipca :607-627, calc_ipk :630-653, calc_ibna :656-675, calc_ibca :678-697
        const double ipca = c.gpca * cai / (c.KpCa + cai);
        const double rec_ipK = E::dv(1., 1. + E::eu(FWB_EDIVK(25 - u, 5.98)));
        const double ipk = c.gpk * rec_ipK * (u - Ek);
        const double ibna = c.gbna * (u - Ena);
        const double ibca = c.gbca * (u - Eca);

        // calc_nai :928-953, calc_ki :956-984 (old nai, Ki)
        {
            const double dNai = -(ina + ibna + 3 * inak + 3 * inaca) * c.inverseVcF * c.CAPACITANCE;
            io.st(3, nai + dt * dNai);
            const double dKi = -(ik1 + ito + ikr + iks - 2 * inak + ipk) * c.inverseVcF * c.CAPACITANCE;
            io.st(4, Ki + dt * dKi);
        }
        un -= dt * (ikr + iks + ik1 + ito + ina + ibna + ical + ibca + inak + inaca + ipca + ipk);

        // calc_irel :700-730
        const double casr = io.ld(1);
        double irel;
        {
            double rr = io.ld(17);
            const double kCaSR = c.maxsr - (c.maxsr_m_minsr / (1 + (c.EC / casr) * (c.EC / casr)));
            const double k1 = c.k1_ / kCaSR;
            const double k2 = c.k2_ * kCaSR;
            const double drr = c.k4 * (1 - rr) - k2 * cass * rr;
            rr += dt * drr;
            const double oo = k1 * cass * cass * rr / (c.k3 + k1 * cass * cass);
            irel = c.Vrel * oo * (casr - cass);
            io.st(17, rr); io.st(18, oo);
        }
        // calc_ileak :733-752, calc_iup :755-774, calc_ixfer :777-796
        const double ileak = c.Vleak * (casr - cai);
        const double iup = c.Vmaxup / (1. + (c.Kup2 / (cai * cai)));
        const double ixfer = c.Vxfer * (cass - cai);
        // calc_casr :799-828
        {
            const double CaCSQN = c.Bufsr * casr / (casr + c.Kbufsr);
            const double dCaSR = dt * (iup - irel - ileak);
            const double bjsr = c.Bufsr - CaCSQN - dCaSR - casr + c.Kbufsr;
            const double cjsr = c.Kbufsr * (CaCSQN + dCaSR + casr);
            io.st(1, (sqrt(bjsr * bjsr + 4 * cjsr) - bjsr) * 0.5);   // == / 2
        }
        // calc_cass :831-872
        {
            const double CaSSBuf = c.Bufss * cass / (cass + c.Kbufss);
            const double dCaSS = dt * (-ixfer * c.Vc_Vss + irel * c.Vsr_Vss +
                                       (-ical * c.inversevssF2 * c.CAPACITANCE));
            const double bcss = c.Bufss - CaSSBuf - dCaSS - cass + c.Kbufss;
            const double ccss = c.Kbufss * (CaSSBuf + dCaSS + cass);
            io.st(2, (sqrt(bcss * bcss + 4 * ccss) - bcss) * 0.5);   // == / 2
        }
        // calc_cai :875-925 is dead: cai keeps its old value (:1098)
    }
};

// ---------------------------------------------------------------------------
// Courtemanche 1998 human atrial model -- cpuwave2D/model/courtemanche_2d.py:180-567
// (point functions), :569-641 (kernel); courtemanche_3d.py:60-163.  SURVEY 8f row f1.
// state slots (model.state_vars without u):
//   0 nai 1 ki 2 cai 3 caup 4 carel 5 m 6 h 7 j_ 8 d 9 f 10 oa 11 oi 12 ua 13 ui
//   14 xr 15 xs 16 fca 17 irel 18 vrel 19 urel 20 wrel
// Reference quirks kept: the kernel's xs/xr parameters are bound to model.xr/model.xs
// (slot 14 is the I_Ks gate, slot 15 the I_Kr gate); calc_ikr ignores the gkr parameter.
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_COURTEMANCHE> {
#ifndef FWB_COURT_MIN_BLOCKS
#define FWB_COURT_MIN_BLOCKS 3
#endif
    static constexpr int NS = 21, NP = 39, MIN_BLOCKS = FWB_COURT_MIN_BLOCKS;
    static constexpr bool USE_TMA = false;
    static constexpr uint32_t READ_MASK = 0x1fffff, WRITE_MASK = 0x1fffff;
    struct Consts {
        double dt, gna, gnab, gk1, gks, gto, gcal, gcab, gkur_coeff, F, ibk, cao, nao, ko;
        double RT_F, RT_2F, kq10, kmnai, ko_kmko, inakmax, nak_s, inacamax, nao3, ncx_t12, ksatncx;
        double ipcamax, krel, iupmax, kup, Vrel, Vup, Vrel_Vup, Fn_a, Fn_b;
        double trpn_k, kmtrpn, cmdn_k, kmcmdn, csqn_k, kmcsqn;
        DivC RT, caupmax, FVj, FVj2, Vj;
        // fast path only (ionic_fast)
        double l_nao, l_ko, l_cao, F_RT, e_fca, e_urel;
        double k47, k32, k5a, k5b, k17a, k17b, k17c, k85;
        int fast_ok;
        double ec[8];      // fexp's reduction / polynomial constants (fexp_fill_consts)
    };
    static bool derive(const double *p, double dt, Consts &c)
    {
        fexp_fill_consts(c.ec);
        const double F = p[9], T = p[10], R = p[11], Vj = p[13], Vup = p[14], Vrel = p[15],
                     nao = p[18], ko = p[19], kmko = p[23], kmnancx = p[24], kmcancx = p[25];
        c.dt = dt; c.gna = p[0]; c.gnab = p[1]; c.gk1 = p[2]; c.gks = p[4]; c.gto = p[5];
        c.gcal = p[6]; c.gcab = p[7]; c.gkur_coeff = p[8]; c.F = F; c.ibk = p[16];
        c.cao = p[17]; c.nao = nao; c.ko = ko;
        c.RT_F = R * T / F; c.RT_2F = R * T / (2 * F); c.RT = make_divc(R * T);
        c.kq10 = p[38]; c.kmnai = p[22]; c.ko_kmko = ko / (ko + kmko); c.inakmax = p[34];
        c.nak_s = (1 / 7.0) * (exp(nao / 67.3) - 1);
        c.inacamax = p[33]; c.nao3 = nao * (nao * nao);
        c.ncx_t12 = ((kmnancx * (kmnancx * kmnancx)) + c.nao3) * (kmcancx + p[17]);
        c.ksatncx = p[26];
        c.ipcamax = p[35]; c.krel = p[36]; c.iupmax = p[37]; c.kup = p[21];
        c.Vrel = Vrel; c.Vup = Vup; c.Vrel_Vup = Vrel / Vup;
        c.Fn_a = 1e-12 * Vrel; c.Fn_b = (5 * 1e-13) / F;
        c.trpn_k = p[30] * p[28]; c.kmtrpn = p[28]; c.cmdn_k = p[31] * p[27]; c.kmcmdn = p[27];
        c.csqn_k = p[32] * p[29]; c.kmcsqn = p[29];
        c.caupmax = make_divc(p[20]); c.FVj = make_divc(F * Vj); c.FVj2 = make_divc(2 * F * Vj);
        c.Vj = make_divc(Vj);
        c.l_nao = log(nao); c.l_ko = log(ko); c.l_cao = log(p[17]); c.F_RT = F / (R * T);
        c.e_fca = exp(-dt / 2); c.e_urel = exp(-dt / 8);
        c.k47 = exp(-0.1 * 47.13); c.k32 = exp(-0.1 * 32.);
        c.k5a = exp(-14.1 / 5.); c.k5b = exp(7.9 / 5.);
        c.k17a = exp(19.9 / 17.); c.k17b = exp(40. / 17.); c.k17c = exp(82. / 17.);
        c.k85 = exp(-10. / 8.5);
        c.fast_ok = nao > 0 && ko > 0 && p[17] > 0 && nao < 1e300 && ko < 1e300 && p[17] < 1e300;
        return divc_ok(c.RT) && divc_ok(c.caupmax) && divc_ok(c.FVj) && divc_ok(c.FVj2) &&
               divc_ok(c.Vj);
    }
    // calc_gating_variable :187-189
    template <class E> FWB_HD static double gate(double x, double x_inf, double tau_x, double dt)
    {
        return x_inf - (x_inf - x) * E::en(-dt / tau_x);
    }
    static constexpr bool HAS_FAST = true;
    // the rearranged path needs |u| < 300 mV and positive, normal nai / ki / cai (flog,
    // branch-free reciprocals); NaNs fail the comparisons and take the reference statement
    template <class IO> FWB_HD static bool fast_ok(double u, const IO &io, const Consts &c)
    {
        return c.fast_ok && fabs(u) < FAST_MATH_U_LIMIT && conc_ok(io.ld(0)) &&
               conc_ok(io.ld(1)) && conc_ok(io.ld(2));
    }
    template <class IO>
    FWB_HD static void ionic_fastpath(double u, double &un, IO &io, const Consts &c)
    {
        ionic_fast(u, un, io, c, io.ld(0), io.ld(1), io.ld(2));
    }
    template <class IO>
    FWB_HD static void ionic_ref(double u, double &un, IO &io, const Consts &c)
    {
        ionic_impl<IO, LibMath>(u, un, io, c, io.ld(0), io.ld(1), io.ld(2));
    }
    template <class IO>
    FWB_HD static void ionic(double u, double &un, IO &io, const Consts &c)
    {
#ifdef __CUDA_ARCH__
        if (fast_ok(u, io, c)) ionic_fastpath(u, un, io, c);
        else
#endif
            ionic_ref(u, un, io, c);
    }
    FWB_HD static bool conc_ok(double x) { return x > 1e-300 && x < 1e300; }
    FWB_HD static double rlf(double inf, double x, double e) { return fma(x - inf, e, inf); }

    // ------------------------------------------------------------------------------------
    // The device's normal path: ionic_impl's equations rearranged for the FP64 pipe, the
    // recipe of TP06 / LR91 (identities only; tests/test_host_models.py compares the two):
    // a Rush-Larsen factor needs 1 / tau, and tau = 1 / (a + b) here, so most gates need no
    // division for it; x_inf = a / (a + b) and the k / (c + exp) rates are put over common
    // denominators (one frcp3 each); exponentials with slopes -1/5, +-1/17, -2/17, -0.1
    // share evaluations; exp(-dt / 2), exp(-dt / 8) (constant time constants) come from the
    // host; logs by flog; x^1.5 = x sqrt(x).
    // ------------------------------------------------------------------------------------
    template <class IO>
    FWB_HD static void ionic_fast(double u, double &un, IO &io, const Consts &c, double nai,
                                  double ki, double cai)
    {
        const double dt = c.dt;
        const double *ec = c.ec;       // kernel-parameter bank -> uniform registers
        auto EX = [ec](double x) { return fexp_fast_p(x, ec); };
        auto EXN = [ec](double x) { return fexp_fast_p(x < -700.0 ? -700.0 : x, ec); };
        auto EXC = [ec](double x) {
            x = x < -700.0 ? -700.0 : x;
            return fexp_fast_p(x > 700.0 ? 700.0 : x, ec);
        };
        (void)EXN; (void)EXC;
        const double ena = c.RT_F * (c.l_nao - flog(nai));
        const double ek = c.RT_F * (c.l_ko - flog(ki));
        const double eca = c.RT_2F * (c.l_cao - flog(cai));
        const double g01 = EX(-0.1 * u);
        const double g5 = EX(-0.2 * u);                    // exp(-u / 5)
        const double g17 = EX(u * (1. / 17.)), i17 = frcp3(g17);
        // ---- I_Na
        double ina;
        {
            const double am = u == -47.13 ? 3.2 : 0.32 * (u + 47.13) * frcp3(fma(-c.k47, g01, 1.));
            const double sm = am + 0.08 * EX(u * (-1. / 11.));
            const double m = rlf(am * frcp3(sm), io.ld(5), EXN(-dt * sm));
            double h_inf, sh, j_inf, sj;
            if (u >= -40) {
                h_inf = 0.;
                sh = frcp3(0.13 * (1. + EX((u + 10.66) * (-1. / 11.1))));
                j_inf = 0.;
                sj = 0.3 * EX(-0.0000002535 * u) * frcp3(fma(c.k32, g01, 1.));
            } else {
                const double ah = 0.135 * EX((80. + u) * (-1. / 6.8));
                sh = ah + (3.56 * EX(0.079 * u) + 310000. * EX(0.35 * u));
                h_inf = ah * frcp3(sh);
                const double a = 1. + EX(0.311 * (u + 79.23));
                const double b = 1. + EX(-0.1378 * (u + 40.14));
                const double nab = (-127140. * EX(0.2444 * u) -
                                    0.00003474 * EX(-0.04391 * u)) * (u + 37.78) * b;
                const double tot = fma(0.1212 * EX(-0.01052 * u), a, nab);   // na b + nb a
                sj = tot * frcp3(a * b);
                j_inf = nab * frcp3(tot);
            }
            const double h = rlf(h_inf, io.ld(6), EXN(-dt * sh));
            const double j = rlf(j_inf, io.ld(7), EXN(-dt * sj));
            io.st(5, m); io.st(6, h); io.st(7, j);
            ina = c.gna * (m * (m * m)) * h * j * (u - ena);
        }
        const double ik1 = c.gk1 * (u - ek) * frcp3(1. + EX(0.07 * (u + 80.)));
        // ---- I_to, I_Kur (share tau_o)
        double ito, ikur;
        {
            const double A = fma(c.k85, i17 * i17, EX((u - 30.) * (-1. / 59.0)));
            const double B = fma(c.k17c, g17, 2.5);
            const double rtau_o = c.kq10 * 0.65 * (A + B) * frcp3(A * B);
            const double o_inf = frcp3(1. + EX((u + 20.47) * (-1. / 17.54)));
            const double C = 18.53 + EX((u + 113.7) * (1. / 10.95));
            const double D = 35.56 + EX((u + 1.26) * (-1. / 7.44));
            const double rtau_oi = c.kq10 * (C + D) * frcp3(C * D);
            const double oi_inf = frcp3(1. + EX((u + 43.1) * (1. / 5.3)));
            const double e_o = EXN(-dt * rtau_o);
            const double oa = rlf(o_inf, io.ld(10), e_o);
            const double oi = rlf(oi_inf, io.ld(11), EXN(-dt * rtau_oi));
            io.st(10, oa); io.st(11, oi);
            ito = c.gto * (oa * (oa * oa)) * oi * (u - ek);

            const double gkur = fma(0.05, frcp3(1. + EX((u - 15.) * (-1. / 13.0))), 0.005);
            const double ua_inf = frcp3(1. + EX((u + 30.3) * (-1. / 9.6)));
            const double rtau_ui = c.kq10 * (frcp3(21. + EX((u - 185.) * (-1. / 28.0))) +
                                             EX((u - 158.) * (1. / 16.0)));
            const double ui_inf = frcp3(1. + EX((u - 99.45) * (1. / 27.48)));
            const double ua = rlf(ua_inf, io.ld(12), e_o);          // tau_ua == tau_o
            const double ui = rlf(ui_inf, io.ld(13), EXN(-dt * rtau_ui));
            io.st(12, ua); io.st(13, ui);
            ikur = c.gkur_coeff * gkur * (ua * (ua * ua)) * ui * (u - ek);
        }
        // ---- I_Kr (gate in slot 15)
        double ikr;
        {
            const double d1 = fma(-c.k5a, g5, 1.);                                   // 1 - exp(-(u+14.1)/5)
            const double d2 = EX((u - 3.3328) * (1. / 5.1237)) - 1.;
            const double rtau_xr = fma(0.0003 * (u + 14.1), d2, 0.000073898 * (u - 3.3328) * d1) *
                                   frcp3(d1 * d2);
            const double xr_inf = frcp3(1. + EX((u + 14.1) * (-1. / 6.5)));
            const double xr = rlf(xr_inf, io.ld(15), EXN(-dt * rtau_xr));
            io.st(15, xr);
            ikr = 0.0294 * xr * (u - ek) * frcp3(1. + EX((u + 15.) * (1. / 22.4)));
        }
        // ---- I_Ks (gate in slot 14)
        double iks;
        {
            const double d1 = fma(-c.k17a, i17, 1.);                                 // 1 - exp(-(u-19.9)/17)
            const double d2 = EX((u - 19.9) * (1. / 9.)) - 1.;
            const double rtau_xs = 2. * (u - 19.9) * fma(0.00004, d2, 0.000035 * d1) * frcp3(d1 * d2);
            const double xs_inf = frcp3(fsqrt(1. + EX((u - 19.9) * (-1. / 12.7))));
            const double xs = rlf(xs_inf, io.ld(14), EXN(-dt * rtau_xs));
            io.st(14, xs);
            iks = c.gks * (xs * xs) * (u - ek);
        }
        // ---- I_CaL
        double ical;
        {
            const double e10 = EX((u + 10.) * (-1. / 6.24));
            const double rtau_d = 0.035 * (u + 10.) * (1. + e10) * frcp3(1. - e10);
            const double d_inf = frcp3(1. + EX((u + 10.) * (-1. / 8.0)));
            const double rtau_f = fma(0.0197, EX(-(0.0337 * 0.0337) * ((u + 10.) * (u + 10.))),
                                      0.02) * (1. / 9.);
            const double f_inf = frcp3(1. + EX((u + 28.) * (1. / 6.9)));
            const double fca_inf = frcp3(fma(cai, 1. / 0.00035, 1.));
            const double d = rlf(d_inf, io.ld(8), EXN(-dt * rtau_d));
            const double f = rlf(f_inf, io.ld(9), EXN(-dt * rtau_f));
            const double fca = rlf(fca_inf, io.ld(16), c.e_fca);
            io.st(8, d); io.st(9, f); io.st(16, fca);
            ical = c.gcal * d * f * fca * (u - 65.);
        }
        // ---- I_NaK, I_NaCa
        const double x = u * c.F_RT;
        const double en01 = EX(-0.1 * x);
        const double en02 = en01 * en01, en04 = en02 * en02, en08 = en04 * en04;
        const double en_x = en08 * en02;                                             // exp(-x)
        const double qn = c.kmnai * frcp3(nai);
        const double inak = c.inakmax * c.ko_kmko *
                            frcp3(fma(0.0365 * c.nak_s, en_x, fma(0.1245, en01, 1.)) *
                                  fma(qn, fsqrt(qn), 1.));
        const double e_rev = EX(-0.65 * x);
        const double inaca = c.inacamax * e_rev * fma(-c.nao3 * cai, en_x, (nai * (nai * nai)) * c.cao) *
                             frcp3(en_x * c.ncx_t12 * fma(c.ksatncx, e_rev, 1.));
        const double ibca = c.gcab * (u - eca);
        const double ibna = c.gnab * (u - ena);
        const double ipca = c.ipcamax * cai * frcp3(cai + 0.0005);
        un -= dt * (ina + ik1 + ito + ikur + ikr + iks + ical + ipca + inak + inaca + ibna + ibca);
        io.st(0, fma(dt, (-3 * inak - 3 * inaca - ibna - ina) * c.FVj.rc, nai));
        io.st(1, fma(dt, (2 * inak - ik1 - ito - ikur - ikr - iks - c.ibk) * c.FVj.rc, ki));
        // ---- SR fluxes
        const double caup = io.ld(3), carel = io.ld(4);
        double irel;
        {
            const double Fn = c.Fn_a * io.ld(17) - c.Fn_b * (0.5 * ical - 0.2 * inaca);
            const double u_inf = frcp3(1. + EXC((Fn - 3.4175e-13) * (-1. / 13.67e-16)));
            const double rtau_v = frcp3(fma(2.09, u_inf, 1.91));
            const double v_inf = 1. - frcp3(1. + EXC((Fn - 6.835e-14) * (-1. / 13.67e-16)));
            const double e79 = c.k5b * g5;                                           // exp(-(u-7.9)/5)
            const double rtau_w = fma(0.3, e79, 1.) * (u - 7.9) * frcp3(6. * (1. - e79));
            const double w_inf = 1. - frcp3(fma(c.k17b, i17, 1.));
            const double urel = rlf(u_inf, io.ld(19), c.e_urel);
            const double vrel = rlf(v_inf, io.ld(18), EXN(-dt * rtau_v));
            const double wrel = rlf(w_inf, io.ld(20), EXN(-dt * rtau_w));
            irel = c.krel * (urel * urel) * vrel * wrel * (carel - cai);
            io.st(17, irel); io.st(19, urel); io.st(18, vrel); io.st(20, wrel);
        }
        const double itr = (caup - carel) * (1. / 180.);
        const double iup = c.iupmax * cai * frcp3(cai + c.kup);
        const double iupleak = caup * c.caupmax.rc * c.iupmax;
        io.st(3, fma(dt, iup - iupleak - itr * c.Vrel_Vup, caup));
        {
            const double B1 = (2 * inaca - ipca - ical - ibca) * c.FVj2.rc +
                              (c.Vup * (iupleak - iup) + irel * c.Vrel) * c.Vj.rc;
            const double P = (cai + c.kmtrpn) * (cai + c.kmtrpn), Q = (cai + c.kmcmdn) * (cai + c.kmcmdn);
            const double PQ = P * Q;
            io.st(2, fma(dt, B1 * PQ * frcp3(fma(c.trpn_k, Q, fma(c.cmdn_k, P, PQ))), cai));
        }
        {
            const double S = (carel + c.kmcsqn) * (carel + c.kmcsqn);
            io.st(4, fma(dt, (itr - irel) * S * frcp3(S + c.csqn_k), carel));
        }
    }

    template <class IO, class E>
    FWB_HD static void ionic_impl(double u, double &un, IO &io, const Consts &c, double nai,
                                  double ki, double cai)
    {
        const double dt = c.dt;
        // calc_equilibrum_potentials :246-251
        const double ena = c.RT_F * log(c.nao / nai);
        const double ek = c.RT_F * log(c.ko / ki);
        const double eca = c.RT_2F * log(c.cao / cai);
        // calc_gating_m :259-271, calc_gating_h :274-285, calc_gating_j :288-299, calc_ina :254
        double ina;
        {
            double am;
            if (u == -47.13) am = 3.2;
            else am = 0.32 * (u + 47.13) / (1 - E::eu(-0.1 * (u + 47.13)));
            const double bm = 0.08 * E::eu(FWB_EDIVK(-u, 11.));
            const double m = gate<E>(io.ld(5), am / (am + bm), 1 / (am + bm), dt);
            double ah, bh, aj, bj;
            if (u >= -40) {
                ah = 0;
                bh = E::dv(1., 0.13 * (1 + E::eu(FWB_EDIVK(-(u + 10.66), 11.1))));
                aj = 0;
                bj = E::dv(0.3 * E::eu(-0.0000002535 * u), 1 + E::eu(-0.1 * (u + 32)));
            } else {
                ah = 0.135 * E::eu(FWB_EDIVK(-(80 + u), 6.8));
                bh = 3.56 * E::eu(0.079 * u) + 310000 * E::eu(0.35 * u);
                aj = E::dv((-127140 * E::eu(0.2444 * u) - 0.00003474 * E::eu(-0.04391 * u)) *
                               (u + 37.78),
                           1 + E::eu(0.311 * (u + 79.23)));
                bj = E::dv(0.1212 * E::eu(-0.01052 * u), 1 + E::eu(-0.1378 * (u + 40.14)));
            }
            const double h = gate<E>(io.ld(6), ah / (ah + bh), 1 / (ah + bh), dt);
            const double j_inf = aj / (aj + bj), tau_j = 1 / (aj + bj);
            const double j = j_inf - (j_inf - io.ld(7)) * E::en(-dt / tau_j);
            io.st(5, m); io.st(6, h); io.st(7, j);
            ina = c.gna * (m * (m * m)) * h * j * (u - ena);
        }
        const double ik1 = c.gk1 * (u - ek) / (1 + E::eu(0.07 * (u + 80)));   // :302-304
        // calc_ito :307-325 and calc_ikur :328-347 (share ao/bo)
        double ito, ikur;
        {
            const double ao = E::dv(0.65, E::eu(FWB_EDIVK(-(u + 10), 8.5)) + E::eu(FWB_EDIVK(-(u - 30), 59.0)));
            const double bo = E::dv(0.65, 2.5 + E::eu(FWB_EDIVK(u + 82, 17.0)));
            const double tau_o = 1 / (c.kq10 * (ao + bo));
            const double o_inf = E::dv(1., 1 + E::eu(FWB_EDIVK(-(u + 20.47), 17.54)));
            const double aoi = E::dv(1., 18.53 + E::eu(FWB_EDIVK(u + 113.7, 10.95)));
            const double boi = E::dv(1., 35.56 + E::eu(FWB_EDIVK(-(u + 1.26), 7.44)));
            const double tau_oi = 1 / (c.kq10 * (aoi + boi));
            const double oi_inf = E::dv(1., 1 + E::eu(FWB_EDIVK(u + 43.1, 5.3)));
            const double oa = gate<E>(io.ld(10), o_inf, tau_o, dt);
            const double oi = gate<E>(io.ld(11), oi_inf, tau_oi, dt);
            io.st(10, oa); io.st(11, oi);
            ito = c.gto * (oa * (oa * oa)) * oi * (u - ek);

            const double gkur = 0.005 + E::dv(0.05, 1 + E::eu(FWB_EDIVK(-(u - 15), 13.0)));
            const double tau_ua = tau_o;           // aua == ao, bua == bo (:330-332)
            const double ua_inf = E::dv(1., 1 + E::eu(FWB_EDIVK(-(u + 30.3), 9.6)));
            const double aui = E::dv(1., 21 + E::eu(FWB_EDIVK(-(u - 185), 28.0)));
            const double bui = E::eu(FWB_EDIVK(u - 158, 16.0));
            const double tau_ui = 1 / (c.kq10 * (aui + bui));
            const double ui_inf = E::dv(1., 1 + E::eu(FWB_EDIVK(u - 99.45, 27.48)));
            const double ua = gate<E>(io.ld(12), ua_inf, tau_ua, dt);
            const double ui = gate<E>(io.ld(13), ui_inf, tau_ui, dt);
            io.st(12, ua); io.st(13, ui);
            ikur = c.gkur_coeff * gkur * (ua * (ua * ua)) * ui * (u - ek);
        }
        // calc_ikr :350-362 -- gate stored in slot 15 (model.xs), gkr fixed at 0.0294
        double ikr;
        {
            const double axr = 0.0003 * (u + 14.1) / (1 - E::eu(FWB_EDIVK(-(u + 14.1), 5.)));
            const double bxr = 0.000073898 * (u - 3.3328) / (E::eu(FWB_EDIVK(u - 3.3328, 5.1237)) - 1);
            const double tau_xr = 1 / (axr + bxr);
            const double xr_inf = E::dv(1., 1 + E::eu(FWB_EDIVK(-(u + 14.1), 6.5)));
            const double xr = gate<E>(io.ld(15), xr_inf, tau_xr, dt);
            io.st(15, xr);
            ikr = E::dv(0.0294 * xr * (u - ek), 1 + E::eu(FWB_EDIVK(u + 15, 22.4)));
        }
        // calc_iks :365-376 -- gate stored in slot 14 (model.xr)
        double iks;
        {
            const double axs = 0.00004 * (u - 19.9) / (1 - E::eu(FWB_EDIVK(-(u - 19.9), 17.)));
            const double bxs = 0.000035 * (u - 19.9) / (E::eu(FWB_EDIVK(u - 19.9, 9.)) - 1);
            const double tau_xs = 1 / (2 * (axs + bxs));
            const double xs_inf = 1 / sqrt(1 + E::eu(FWB_EDIVK(-(u - 19.9), 12.7)));
            const double xs = gate<E>(io.ld(14), xs_inf, tau_xs, dt);
            io.st(14, xs);
            iks = c.gks * (xs * xs) * (u - ek);
        }
        // calc_ical :379-396
        double ical;
        {
            const double e10 = E::eu(FWB_EDIVK(-(u + 10), 6.24));
            const double tau_d = (1 - e10) / (0.035 * (u + 10) * (1 + e10));
            const double d_inf = E::dv(1., 1 + E::eu(FWB_EDIVK(-(u + 10), 8.0)));
            const double tau_f = 9 / (0.0197 * E::eu(-(0.0337 * 0.0337) * ((u + 10) * (u + 10))) + 0.02);
            const double f_inf = E::dv(1., 1 + E::eu(FWB_EDIVK(u + 28, 6.9)));
            const double fca_inf = 1 / (1 + FWB_EDIVK(cai, 0.00035));
            const double d = gate<E>(io.ld(8), d_inf, tau_d, dt);
            const double f = gate<E>(io.ld(9), f_inf, tau_f, dt);
            const double fca = gate<E>(io.ld(16), fca_inf, 2, dt);
            io.st(8, d); io.st(9, f); io.st(16, fca);
            ical = c.gcal * d * f * fca * (u - 65);
        }
        // calc_inak :399-404, calc_inaca :407-423
        const double Fu = c.F * u;
        const double fnak = 1 / (1 + 0.1245 * E::eu(E::dk(-0.1 * Fu, c.RT)) +
                                 0.0365 * c.nak_s * E::eu(E::dk(-Fu, c.RT)));
        const double inak = c.inakmax * fnak * (1 / (1 + E::p15(c.kmnai / nai))) * c.ko_kmko;
        double inaca;
        {
            const double exp_term = E::eu(E::dk(0.35 * Fu, c.RT));
            const double exp_rev_term = E::eu(E::dk((0.35 - 1) * Fu, c.RT));
            const double numerator = c.inacamax * (exp_term * (nai * (nai * nai)) * c.cao -
                                                   exp_rev_term * c.nao3 * cai);
            inaca = numerator / (c.ncx_t12 * (1 + c.ksatncx * exp_rev_term));
        }
        const double ibca = c.gcab * (u - eca);               // :426-428
        const double ibna = c.gnab * (u - ena);               // :431-433
        const double ipca = c.ipcamax * cai / (cai + 0.0005); // :436-438
        // membrane potential and the two concentrations that need no SR fluxes
        un -= dt * (ina + ik1 + ito + ikur + ikr + iks + ical + ipca + inak + inaca + ibna + ibca);
        io.st(0, nai + dt * E::dk(-3 * inak - 3 * inaca - ibna - ina, c.FVj));               // :207-210
        io.st(1, ki + dt * E::dk(2 * inak - ik1 - ito - ikur - ikr - iks - c.ibk, c.FVj));    // :213-216
        // calc_irel :441-462
        const double caup = io.ld(3), carel = io.ld(4);
        double irel;
        {
            const double Fn = c.Fn_a * io.ld(17) - c.Fn_b * (0.5 * ical - 0.2 * inaca);
            const double eFn = E::ec(FWB_EDIVK(-(Fn - 3.4175e-13), 13.67e-16));
            const double u_inf = 1 / (1 + eFn);
            const double tau_v = 1.91 + 2.09 / (1 + eFn);
            const double v_inf = 1 - 1 / (1 + E::ec(FWB_EDIVK(-(Fn - 6.835e-14), 13.67e-16)));
            const double e79 = E::eu(FWB_EDIVK(-(u - 7.9), 5.0));
            const double tau_w = 6 * (1 - e79) / ((1 + 0.3 * e79) * (u - 7.9));
            const double w_inf = 1 - E::dv(1., 1 + E::eu(FWB_EDIVK(-(u - 40), 17.0)));
            const double urel = gate<E>(io.ld(19), u_inf, 8, dt);
            const double vrel = gate<E>(io.ld(18), v_inf, tau_v, dt);
            const double wrel = gate<E>(io.ld(20), w_inf, tau_w, dt);
            irel = c.krel * (urel * urel) * vrel * wrel * (carel - cai);
            io.st(17, irel); io.st(19, urel); io.st(18, vrel); io.st(20, wrel);
        }
        const double itr = FWB_EDIVK(caup - carel, 180.);       // :465-468
        const double iup = c.iupmax / (1 + (c.kup / cai));     // :471-473
        const double iupleak = E::dk(caup, c.caupmax) * c.iupmax;   // :476-478
        io.st(3, caup + dt * (iup - iupleak - itr * c.Vrel_Vup));                           // :226-229
        {
            const double B1 = E::dk(2 * inaca - ipca - ical - ibca, c.FVj2) +
                              E::dk(c.Vup * (iupleak - iup) + irel * c.Vrel, c.Vj);
            const double B2 = 1 + c.trpn_k / ((cai + c.kmtrpn) * (cai + c.kmtrpn)) +
                              c.cmdn_k / ((cai + c.kmcmdn) * (cai + c.kmcmdn));
            io.st(2, cai + dt * (B1 / B2));                                                  // :219-224
        }
        io.st(4, carel + dt * ((itr - irel) /
                               (1 + c.csqn_k / ((carel + c.kmcsqn) * (carel + c.kmcsqn)))));  // :232-235
    }
};

}  // namespace fwb
