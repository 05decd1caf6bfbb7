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
//                     leaves as u_new; s[] are the node's state values.
//
// The arithmetic keeps the reference's association order; the translation
// unit is compiled with -fmad=false so no multiply-add is contracted.
// Reference line numbers are cited per model.
#pragma once
#include <math.h>

#include "fwb_common.cuh"

#ifdef __CUDACC__
#define FWB_HD __host__ __device__ __forceinline__
#else
#define FWB_HD inline
#endif

namespace fwb {

template <int MODEL> struct Model;

// ---------------------------------------------------------------------------
// Aliev-Panfilov -- finitewave/cpuwave2D/model/aliev_panfilov_2d.py:133-170
// (calc_v), :173-202 (ionic_kernel_2d); 3D: cpuwave3D/model/aliev_panfilov_3d.py:53-84
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_ALIEV_PANFILOV> {
    static constexpr int NS = 1, NP = 5;
    static constexpr uint32_t READ_MASK = 0x1, WRITE_MASK = 0x1;
    struct Consts { double dt, a, k, eap, mu_1, mu_2; };
    static void derive(const double *p, double dt, Consts &c)
    {
        c.dt = dt; c.a = p[0]; c.k = p[1]; c.eap = p[2]; c.mu_1 = p[3]; c.mu_2 = p[4];
    }
    FWB_HD static void ionic(double u, double &un, double *s, const Consts &c)
    {
        double v = s[0];
        v += (-c.dt * (c.eap + (c.mu_1 * v) / (c.mu_2 + u)) * (v + c.k * u * (u - c.a - 1.)));
        s[0] = v;
        un += c.dt * (-c.k * u * (u - c.a) * (u - 1.) - u * v);
    }
};

// ---------------------------------------------------------------------------
// Barkley -- cpuwave2D/model/barkley_2d.py:113-137, :140-166; barkley_3d.py:53-83
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_BARKLEY> {
    static constexpr int NS = 1, NP = 3;
    static constexpr uint32_t READ_MASK = 0x1, WRITE_MASK = 0x1;
    struct Consts { double dt, a, b, eap; };
    static void derive(const double *p, double dt, Consts &c)
    {
        c.dt = dt; c.a = p[0]; c.b = p[1]; c.eap = p[2];
    }
    FWB_HD static void ionic(double u, double &un, double *s, const Consts &c)
    {
        double v = s[0];
        v += c.dt * (u - v);
        s[0] = v;
        un += c.dt * (u * (1 - u) * (u - (v + c.b) / c.a)) / c.eap;
    }
};

// ---------------------------------------------------------------------------
// Mitchell-Schaeffer -- cpuwave2D/model/mitchell_schaeffer_2d.py:102-186,
// :188-217; mitchell_schaeffer_3d.py:57-95.  strict `u < u_gate`.
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_MITCHELL_SCHAEFFER> {
    static constexpr int NS = 1, NP = 5;
    static constexpr uint32_t READ_MASK = 0x1, WRITE_MASK = 0x1;
    struct Consts { double dt, tau_close, tau_open, tau_in, tau_out, u_gate; };
    static void derive(const double *p, double dt, Consts &c)
    {
        c.dt = dt; c.tau_close = p[0]; c.tau_open = p[1]; c.tau_in = p[2];
        c.tau_out = p[3]; c.u_gate = p[4];
    }
    FWB_HD static void ionic(double u, double &un, double *s, const Consts &c)
    {
        double h = s[0];
        h += (u < c.u_gate) ? c.dt * (1.0 - h) / c.tau_open : c.dt * (-h) / c.tau_close;
        s[0] = h;
        const double C = (u * u) * (1 - u);
        const double J_in = h * C / c.tau_in;
        const double J_out = -u / c.tau_out;
        un += c.dt * (J_in + J_out);
    }
};

// ---------------------------------------------------------------------------
// Fenton-Karma -- cpuwave2D/model/fenton_karma_2d.py:143-292, :294-328;
// fenton_karma_3d.py:61-97.  Both Heavisides are 1 at u == u_c; J_si is
// evaluated with v (not w) exactly as the reference does (:326).
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_FENTON_KARMA> {
    static constexpr int NS = 2, NP = 11;
    static constexpr uint32_t READ_MASK = 0x3, WRITE_MASK = 0x3;
    struct Consts {
        double dt, tau_d, tau_o, tau_r, tau_si, tau_v_m, tau_v_p, tau_w_m, tau_w_p, k,
            u_c, uc_si, two_tau_si;
    };
    static void derive(const double *p, double dt, Consts &c)
    {
        c.dt = dt; c.tau_d = p[0]; c.tau_o = p[1]; c.tau_r = p[2]; c.tau_si = p[3];
        c.tau_v_m = p[4]; c.tau_v_p = p[5]; c.tau_w_m = p[6]; c.tau_w_p = p[7];
        c.k = p[8]; c.u_c = p[9]; c.uc_si = p[10];
        c.two_tau_si = 2 * c.tau_si;
    }
    FWB_HD static void ionic(double u, double &un, double *s, const Consts &c)
    {
        const double H1 = (c.u_c - u >= 0) ? 1.0 : 0.0;
        const double H2 = (u - c.u_c >= 0) ? 1.0 : 0.0;
        double v = s[0], w = s[1];
        v += c.dt * (H1 * (1 - v) / c.tau_v_m - H2 * v / c.tau_v_p);
        w += c.dt * (H1 * (1 - w) / c.tau_w_m - H2 * w / c.tau_w_p);
        s[0] = v;
        s[1] = w;
        const double J_fi = -(v * H2 * (1 - u) * (u - c.u_c)) / c.tau_d;
        const double J_so = u * H1 / c.tau_o + H2 / c.tau_r;
        const double J_si = -v * (1 + tanh(c.k * (u - c.uc_si))) / c.two_tau_si;
        un += c.dt * (-J_fi - J_so - J_si);
    }
};

// ---------------------------------------------------------------------------
// Luo-Rudy 1991 -- cpuwave2D/model/luo_rudy91_2d.py:158-443, :446-510;
// luo_rudy91_3d.py:63-132.  Gates are forward Euler (calc_gating_var :158-182),
// I_Na uses the new m,h,j, I_si the old d,f, I_K the old x.
// state: m,h,j,d,f,x,cai
// ---------------------------------------------------------------------------
template <> struct Model<FWB_MODEL_LUO_RUDY91> {
    static constexpr int NS = 7, NP = 15;
    static constexpr uint32_t READ_MASK = 0x7f, WRITE_MASK = 0x7f;
    struct Consts {
        double dt, gna, gsi, gkp, gb;
        double E_Na, E_K, G_K, E_K1, G_K1;   // parameter-only (:478, :337-338, :496, :404)
    };
    static void derive(const double *p, double dt, Consts &c)
    {
        const double gk = p[2], gk1 = p[3], ko = p[6], ki = p[7], nai = p[8], nao = p[9],
                     R = p[11], T = p[12], F = p[13], PR_NaK = p[14];
        c.dt = dt; c.gna = p[0]; c.gsi = p[1]; c.gkp = p[4]; c.gb = p[5];
        c.E_Na = (R * T / F) * log(nao / nai);
        c.E_K = (R * T / F) * log((ko + PR_NaK * nao) / (ki + PR_NaK * nai));
        c.G_K = gk * sqrt(ko / 5.4);
        c.E_K1 = (R * T / F) * log(ko / ki);
        c.G_K1 = gk1 * sqrt(ko / 5.4);
    }
    FWB_HD static double gate(double var, double dt, double alpha, double beta)
    {
        const double tau = 1. / (alpha + beta);
        const double inf = alpha / (alpha + beta);
        var += dt * (inf - var) / tau;
        return var;
    }
    FWB_HD static void ionic(double u, double &un, double *s, const Consts &c)
    {
        const double dt = c.dt;
        // calc_ina :185-241
        double alpha_h = 0, beta_h = 0, beta_J = 0, alpha_J = 0;
        if (u >= -40.) {
            beta_h = 1. / (0.13 * (1 + exp((u + 10.66) / -11.1)));
            beta_J = 0.3 * exp(-2.535 * 1e-07 * u) / (1 + exp(-0.1 * (u + 32)));
        } else {
            alpha_h = 0.135 * exp((80 + u) / -6.8);
            beta_h = 3.56 * exp(0.079 * u) + 3.1 * 1e5 * exp(0.35 * u);
            beta_J = 0.1212 * exp(-0.01052 * u) / (1 + exp(-0.1378 * (u + 40.14)));
            alpha_J = (-1.2714 * 1e5 * exp(0.2444 * u) - 3.474 * 1e-5 * exp(-0.04391 * u)) *
                      (u + 37.78) / (1 + exp(0.311 * (u + 79.23)));
        }
        const double alpha_m = 0.32 * (u + 47.13) / (1 - exp(-0.1 * (u + 47.13)));
        const double beta_m = 0.08 * exp(-u / 11);
        const double m = gate(s[0], dt, alpha_m, beta_m);
        const double h = gate(s[1], dt, alpha_h, beta_h);
        const double j = gate(s[2], dt, alpha_J, beta_J);
        s[0] = m; s[1] = h; s[2] = j;
        const double ina = c.gna * m * m * m * h * j * (u - c.E_Na);
        // calc_isk :244-294
        double d = s[3], f = s[4], cai = s[6];
        const double E_Si = 7.7 - 13.0287 * log(cai);
        const double I_Si = c.gsi * d * f * (u - E_Si);
        const double alpha_d = 0.095 * exp(-0.01 * (u - 5)) / (1 + exp(-0.072 * (u - 5)));
        const double beta_d = 0.07 * exp(-0.017 * (u + 44)) / (1 + exp(0.05 * (u + 44)));
        const double alpha_f = 0.012 * exp(-0.008 * (u + 28)) / (1 + exp(0.15 * (u + 28)));
        const double beta_f = 0.0065 * exp(-0.02 * (u + 30)) / (1 + exp(-0.2 * (u + 30)));
        d = gate(d, dt, alpha_d, beta_d);
        f = gate(f, dt, alpha_f, beta_f);
        cai += dt * (-0.0001 * I_Si + 0.07 * (0.0001 - cai));
        s[3] = d; s[4] = f; s[6] = cai;
        // calc_ik :297-356
        double Xi;
        if (u > -100)
            Xi = 2.837 * (exp(0.04 * (u + 77)) - 1) / ((u + 77) * exp(0.04 * (u + 35)));
        else
            Xi = 1;
        double x = s[5];
        const double I_K = c.G_K * x * Xi * (u - c.E_K);
        const double alpha_x = 0.0005 * exp(0.083 * (u + 50)) / (1 + exp(0.057 * (u + 50)));
        const double beta_x = 0.0013 * exp(-0.06 * (u + 20)) / (1 + exp(-0.04 * (u + 20)));
        x = gate(x, dt, alpha_x, beta_x);
        s[5] = x;
        // calc_ik1 :359-408, calc_ikp :411-427, calc_ib :430-443, kernel :496-510
        const double E_K1 = c.E_K1;
        const double alpha_K1 = 1.02 / (1 + exp(0.2385 * (u - E_K1 - 59.215)));
        const double beta_K1 = (0.49124 * exp(0.08032 * (u - E_K1 + 5.476)) +
                                exp(0.06175 * (u - E_K1 - 594.31))) /
                               (1 + exp(-0.5143 * (u - E_K1 + 4.753)));
        const double K_1x = alpha_K1 / (alpha_K1 + beta_K1);
        const double ik1 = c.G_K1 * K_1x * (u - E_K1);
        const double K_p = 1. / (1 + exp((7.488 - u) / 5.98));
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
    static constexpr int NS = 19, NP = 49;
    static constexpr uint32_t READ_MASK = 0x3ffff;          // all but oo
    static constexpr uint32_t WRITE_MASK = 0x7ffff & ~0x1u;  // all but cai
    struct Consts {
        double dt, ko, cao, nao, RTONF, half_RTONF, CAPACITANCE;
        double pKNa_nao, ko_pKNa_nao, pKNa;
        double gna, gcal, gto, gkr_sqrt, gks, gk1, gpca, KpCa, gpk, gbna, gbca;
        double F, RT, FF_RT;
        double inaca_pref, ksat, n_, n_m1, knak_pref, KmNa;
        double Bufsr, Kbufsr, Bufss, Kbufss, Vmaxup, Kup2, Vrel, k1_, k2_, k3, k4, EC;
        double maxsr, maxsr_m_minsr, Vleak, Vxfer, Vc_Vss, Vsr_Vss;
        double inverseVcF, inversevssF2;
    };
    static void derive(const double *p, double dt, Consts &c)
    {
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
        c.F = F; c.RT = R * T; c.FF_RT = F * F / (R * T);
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
    }
    FWB_HD static void ionic(double u, double &un, double *s, const Consts &c)
    {
        const double dt = c.dt;
        const double cai = s[0], casr = s[1], cass = s[2], nai = s[3], Ki = s[4];
        // reversal potentials :1072-1075
        const double Ek = c.RTONF * log(c.ko / Ki);
        const double Ena = c.RTONF * log(c.nao / nai);
        const double Eks = c.RTONF * log(c.ko_pKNa_nao / (Ki + c.pKNa * nai));
        const double Eca = c.half_RTONF * log(c.cao / cai);

        // calc_ina :242-318
        double m = s[5], h = s[6], j = s[7];
        {
            const double alpha_m = 1. / (1. + exp((-60. - u) / 5.));
            const double beta_m = 0.1 / (1. + exp((u + 35.) / 5.)) +
                                  0.10 / (1. + exp((u - 50.) / 200.));
            const double tau_m = alpha_m * beta_m;
            const double em = 1. + exp((-56.86 - u) / 9.03);
            const double m_inf = 1. / (em * em);
            double alpha_h, beta_h, alpha_j, beta_j;
            if (u >= -40.) {
                alpha_h = 0.;
                beta_h = 0.77 / (0.13 * (1. + exp(-(u + 10.66) / 11.1)));
                alpha_j = 0.;
                beta_j = 0.6 * exp(0.057 * u) / (1. + exp(-0.1 * (u + 32.)));
            } else {
                alpha_h = 0.057 * exp(-(u + 80.) / 6.8);
                beta_h = 2.7 * exp(0.079 * u) + 3.1e5 * exp(0.3485 * u);
                alpha_j = (-2.5428e4 * exp(0.2444 * u) - 6.948e-6 * exp(-0.04391 * u)) *
                          (u + 37.78) / (1. + exp(0.311 * (u + 79.23)));
                beta_j = 0.02424 * exp(-0.01052 * u) / (1. + exp(-0.1378 * (u + 40.14)));
            }
            const double tau_h = 1.0 / (alpha_h + beta_h);
            const double eh = 1. + exp((u + 71.55) / 7.43);
            const double h_inf = 1. / (eh * eh);
            const double tau_j = 1.0 / (alpha_j + beta_j);
            const double j_inf = h_inf;
            m = m_inf - (m_inf - m) * exp(-dt / tau_m);
            h = h_inf - (h_inf - h) * exp(-dt / tau_h);
            j = j_inf - (j_inf - j) * exp(-dt / tau_j);
        }
        s[5] = m; s[6] = h; s[7] = j;
        const double ina = c.gna * m * m * m * h * j * (u - Ena);

        // calc_ical :321-380
        double d = s[13], f = s[14], f2 = s[15], fcass = s[16];
        double ical;
        {
            const double d_inf = 1. / (1. + exp((-8 - u) / 7.5));
            const double Ad = 1.4 / (1. + exp((-35 - u) / 13)) + 0.25;
            const double Bd = 1.4 / (1. + exp((u + 5) / 5));
            const double Cd = 1. / (1. + exp((50 - u) / 20));
            const double tau_d = Ad * Bd + Cd;
            const double f_inf = 1. / (1. + exp((u + 20) / 7));
            const double Af = 1102.5 * exp(-(u + 27) * (u + 27) / 225);
            const double Bf = 200. / (1 + exp((13 - u) / 10.));
            const double e30 = exp((u + 30) / 10);
            const double Cf = (180. / (1 + e30)) + 20;
            const double tau_f = Af + Bf + Cf;
            const double f2_inf = 0.67 / (1. + exp((u + 35) / 7)) + 0.33;
            const double Af2 = 600 * exp(-(u + 25) * (u + 25) / 170);
            const double Bf2 = 31 / (1. + exp((25 - u) / 10));
            const double Cf2 = 16 / (1. + e30);
            const double tau_f2 = Af2 + Bf2 + Cf2;
            const double cq = 1 + (cass / 0.05) * (cass / 0.05);
            const double fcass_inf = 0.6 / cq + 0.4;
            const double tau_fcass = 80. / cq + 2.;
            d = d_inf - (d_inf - d) * exp(-dt / tau_d);
            f = f_inf - (f_inf - f) * exp(-dt / tau_f);
            f2 = f2_inf - (f2_inf - f2) * exp(-dt / tau_f2);
            fcass = fcass_inf - (fcass_inf - fcass) * exp(-dt / tau_fcass);
            const double e2 = exp(2 * (u - 15) * c.F / c.RT);
            ical = c.gcal * d * f * f2 * fcass * 4 * (u - 15) * c.FF_RT *
                   (0.25 * e2 * cass - c.cao) / (e2 - 1.);
        }
        s[13] = d; s[14] = f; s[15] = f2; s[16] = fcass;

        // calc_ito :383-413
        double r = s[11], sg = s[12];
        {
            const double r_inf = 1. / (1. + exp((20 - u) / 6.));
            const double s_inf = 1. / (1. + exp((u + 20) / 5.));
            const double tau_r = 9.5 * exp(-(u + 40.) * (u + 40.) / 1800.) + 0.8;
            const double tau_s = 85. * exp(-(u + 45.) * (u + 45.) / 320.) +
                                 5. / (1. + exp((u - 20.) / 5.)) + 3.;
            sg = s_inf - (s_inf - sg) * exp(-dt / tau_s);
            r = r_inf - (r_inf - r) * exp(-dt / tau_r);
        }
        s[11] = r; s[12] = sg;
        const double ito = c.gto * r * sg * (u - Ek);

        // calc_ikr :416-452
        double xr1 = s[8], xr2 = s[9];
        {
            const double xr1_inf = 1. / (1. + exp((-26. - u) / 7.));
            const double axr1 = 450. / (1. + exp((-45. - u) / 10.));
            const double bxr1 = 6. / (1. + exp((u - (-30.)) / 11.5));
            const double tau_xr1 = axr1 * bxr1;
            const double xr2_inf = 1. / (1. + exp((u - (-88.)) / 24.));
            const double axr2 = 3. / (1. + exp((-60. - u) / 20.));
            const double bxr2 = 1.12 / (1. + exp((u - 60.) / 20.));
            const double tau_xr2 = axr2 * bxr2;
            xr1 = xr1_inf - (xr1_inf - xr1) * exp(-dt / tau_xr1);
            xr2 = xr2_inf - (xr2_inf - xr2) * exp(-dt / tau_xr2);
        }
        s[8] = xr1; s[9] = xr2;
        const double ikr = c.gkr_sqrt * xr1 * xr2 * (u - Ek);

        // calc_iks :455-485
        double xs = s[10];
        {
            const double xs_inf = 1. / (1. + exp((-5. - u) / 14.));
            const double Axs = (1400. / (sqrt(1. + exp((5. - u) / 6))));
            const double Bxs = (1. / (1. + exp((u - 35.) / 15.)));
            const double tau_xs = Axs * Bxs + 80;
            xs = xs_inf - (xs_inf - xs) * exp(-dt / tau_xs);
        }
        s[10] = xs;
        const double iks = c.gks * xs * xs * (u - Eks);

        // calc_ik1 :488-514
        const double ak1 = 0.1 / (1. + exp(0.06 * (u - Ek - 200)));
        const double bk1 = (3. * exp(0.0002 * (u - Ek + 100)) + exp(0.1 * (u - Ek - 10))) /
                           (1. + exp(-0.5 * (u - Ek)));
        const double rec_iK1 = ak1 / (ak1 + bk1);
        const double ik1 = c.gk1 * rec_iK1 * (u - Ek);
        // calc_inaca :517-565
        const double e_nm1 = exp(c.n_m1 * u * c.F / c.RT);
        const double inaca = c.inaca_pref * (1. / (1 + c.ksat * e_nm1)) *
                             (exp(c.n_ * u * c.F / c.RT) * nai * nai * nai * c.cao -
                              e_nm1 * c.nao * c.nao * c.nao * cai * 2.5);
        // calc_inak :568-604
        const double rec_iNaK = (1. / (1. + 0.1245 * exp(-0.1 * u * c.F / c.RT) +
                                       0.0353 * exp(-u * c.F / c.RT)));
        const double inak = c.knak_pref * (nai / (nai + c.KmNa)) * rec_iNaK;
        // calc_ipca :607-627, calc_ipk :630-653, calc_ibna :656-675, calc_ibca :678-697
        const double ipca = c.gpca * cai / (c.KpCa + cai);
        const double rec_ipK = 1. / (1. + exp((25 - u) / 5.98));
        const double ipk = c.gpk * rec_ipK * (u - Ek);
        const double ibna = c.gbna * (u - Ena);
        const double ibca = c.gbca * (u - Eca);

        // calc_irel :700-730
        double rr = s[17], oo, irel;
        {
            const double kCaSR = c.maxsr - (c.maxsr_m_minsr / (1 + (c.EC / casr) * (c.EC / casr)));
            const double k1 = c.k1_ / kCaSR;
            const double k2 = c.k2_ * kCaSR;
            const double drr = c.k4 * (1 - rr) - k2 * cass * rr;
            rr += dt * drr;
            oo = k1 * cass * cass * rr / (c.k3 + k1 * cass * cass);
            irel = c.Vrel * oo * (casr - cass);
        }
        s[17] = rr; s[18] = oo;
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
            s[1] = (sqrt(bjsr * bjsr + 4 * cjsr) - bjsr) / 2;
        }
        // calc_cass :831-872
        {
            const double CaSSBuf = c.Bufss * cass / (cass + c.Kbufss);
            const double dCaSS = dt * (-ixfer * c.Vc_Vss + irel * c.Vsr_Vss +
                                       (-ical * c.inversevssF2 * c.CAPACITANCE));
            const double bcss = c.Bufss - CaSSBuf - dCaSS - cass + c.Kbufss;
            const double ccss = c.Kbufss * (CaSSBuf + dCaSS + cass);
            s[2] = (sqrt(bcss * bcss + 4 * ccss) - bcss) / 2;
        }
        // calc_cai :875-925 is dead: cai keeps its old value (:1098)
        // calc_nai :928-953, calc_ki :956-984
        {
            const double dNai = -(ina + ibna + 3 * inak + 3 * inaca) * c.inverseVcF * c.CAPACITANCE;
            s[3] = nai + dt * dNai;
            const double dKi = -(ik1 + ito + ikr + iks - 2 * inak + ipk) * c.inverseVcF * c.CAPACITANCE;
            s[4] = Ki + dt * dKi;
        }
        un -= dt * (ikr + iks + ik1 + ito + ina + ibna + ical + ibca + inak + inaca + ipca + ipk);
    }
};

}  // namespace fwb
