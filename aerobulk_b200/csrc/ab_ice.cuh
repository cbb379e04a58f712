// ab_ice.cuh -- sea-ice bulk algorithms as device functions.
// Reference: src/ice/mod_blk_ice_{nemo,easy,an05,lu12,lg15,lg15_io}.f90, src/ice/mod_cdn_form_ice.f90 and the ice
// helpers of src/mod_phymbl.f90.  Not provided: mod_blk_ice_best.f90 (reads sqrtCdn10 before setting it, :154 --
// undefined in the reference) and the over-water outputs of turb_ice_lg15_io (never initialised, :170-172).
#pragma once

#include "ab_device.cuh"

#define ABD __device__ __forceinline__

namespace abd {

enum IceAlgo { ICE_NEMO = 1, ICE_EASY = 2, ICE_AN05 = 3, ICE_LU12 = 4, ICE_LG15 = 5 };   // lg15_io == lg15 over the ice

constexpr double RTT0 = 273.16;             // mod_const.f90:61
constexpr double RLSUB = 2.834e+6;          // :92
constexpr double RCD_ICE = 1.4e-3;          // :118
constexpr double WSPD_THRSHLD_ICE = 0.2;    // :120
constexpr double RICE_ALB0 = 0.8;           // mod_const.f90:51
constexpr double EMISS_I = 0.996;           // mod_const.f90:56
constexpr double RDG_I = 0.7858350313586662;            // LOG10(6.1071)   mod_phymbl.f90:147
constexpr double RZ0_I_S_0 = 0.69e-3, RZ0_I_F_0 = 4.54e-4, RALPHA_0 = 0.2;   // mod_blk_ice_lg15.f90:37-41
constexpr double LOG_5 = 0x1.9c041f7ed8d33p+0;          // LOG(1./0.2)

// Goff over ice, mod_phymbl.f90:815-830; q_sat(l_ice=.TRUE.), :881-904
ABD double e_sat_ice(double T)
{
    const double zta = abm::dmax(T, 180.);
    const double ztmp = fdiv(RTT0, zta);
    const double zle = -9.09718 * (ztmp - 1.) + -3.56654 * abm::dlog10(ztmp) + 0.876793 * (1. - zta * (1. / RTT0)) + RDG_I;
    return 100. * abm::dexp10(zle);
}
ABD double q_sat_ice(double T, double p) { return q_sat_from_e(e_sat_ice(T), p); }

// Louis (1979) stability functions, mod_phymbl.f90:1419-1479 (rc_louis = 5); only the selected side is evaluated
ABD double f_louis(double ra, double zu, double Rib, double Cxn, double z0)
{
    if (nonneg(Rib)) return abm::fast_rcp(1. + ra * fdiv(Rib, abm::fast_sqrt(fabs(1. + Rib))));
    const double ztu = fdiv(Rib, 1. + 3. * 25. * Cxn * abm::fast_sqrt(fabs(-Rib * (fdiv(zu, z0) + 1.))));
    return 1. - ra * ztu;
}
ABD double f_m_louis(double zu, double Rib, double Cdn, double z0) { return f_louis(10., zu, Rib, Cdn, z0); }
ABD double f_h_louis(double zu, double Rib, double Chn, double z0) { return f_louis(15., zu, Rib, Chn, z0); }

// mod_cdn_form_ice.f90:170-208 (LU13, level-4 approximation) and :299-330 (LG15 Eq.46)
ABD double CdN10_f_LU13(double A) { return 2.23E-3 * powr(1. - A, 1. + 1. / (10. * 1.4)); }   // A**(mu-1) == 1 (mu = 1)
ABD double CdN_f_LG15_light(double log_ratio, double A)   // log_ratio = LOG(10/z0w) / LOG(zu/z0w), host-computed
{
    return 3.46e-3 * log_ratio * log_ratio * A * powr(1. - A, 1.4);
}

// psi_m_ice / psi_h_ice, mod_blk_ice_an05.f90:329-405 == mod_blk_ice_easy.f90:213-289 (Paulson with 16 / Holtslag-De Bruin)
ABD double psi_ice_stable(double z) { return -(0.7 * z + 0.75 * (z - 14.3) * abm::dexp_b(-0.35 * z) + 10.7); }   // 0 <= z <= 50
ABD double psi_m_ice(double z)
{
    if (nonneg(z)) return psi_ice_stable(z);
    const double x2 = abm::fast_sqrt(fabs(1. - 16. * z)), x = abm::fast_sqrt(x2);
    return abm::dlog((1. + x2) * 0.5) + 2. * abm::dlog((1. + x) * 0.5) - 2. * abm::datan(x) + 0.5 * RPI;
}
ABD double psi_h_ice(double z)
{
    if (nonneg(z)) return psi_ice_stable(z);
    const double x2 = abm::fast_sqrt(fabs(1. - 16. * z));
    return 2. * abm::dlog((1. + x2) * 0.5);
}

// Andreas et al. 2005 Eq.19, mod_blk_ice_an05.f90:247-268
ABD double rough_leng_m(double us, double nua)
{
    const double zus = abm::dmax(us, 1.E-9);
    const double zz = (zus - 0.18) * 10.;
    return fdiv(0.135 * nua, zus) + 0.035 * zus * zus * INV_GRAV * (5. * abm::dexp(-zz * zz) + 1.);
}
// Andreas 1987 table, :270-325, in log space: LOG(z0t) = LOG(z0) + b0 + b1 LOG(R*) + b2 LOG(R*)**2.
// `bad` is raised where the reference would ctl_stop (:296-297: no regime selected for 2.49999 < R* < 2.5).
ABD void log_rough_leng_tq(double z0, double log_z0, double us, double nua, double &log_z0t, double &log_z0q, bool &bad)
{
    const double zus = abm::dmax(us, 1.E-9);
    const double zre = abm::dmax(fdiv(zus * z0, nua), 0.);
    const double zlog = abm::dlog(zre), zlog2 = zlog * zlog;
    double t0, t1, t2, q0, q1, q2;
    if (zre <= 0.135) {
        t0 = 1.25; t1 = 0.; t2 = 0.; q0 = 1.61; q1 = 0.; q2 = 0.;
    } else if (zre <= 2.49999) {
        t0 = 0.149; t1 = -0.550; t2 = 0.; q0 = 0.351; q1 = -0.628; q2 = 0.;
    } else if (zre >= 2.5) {
        t0 = 0.317; t1 = -0.565; t2 = -0.183; q0 = 0.396; q1 = -0.512; q2 = -0.180;
    } else {
        t0 = t1 = t2 = q0 = q1 = q2 = 0.;
        bad = true;
    }
    log_z0t = log_z0 + (t0 + t1 * zlog + t2 * zlog2);
    log_z0q = log_z0 + (q0 + q1 * zlog + q2 * zlog2);
}

struct IceUniform {
    double zt, zu, log_zu, log_ztu, log_zu10;
    double cxn[3], sqrt_cdn;          // easy: scalar CdN, ChN, CeN inputs and SQRT(CdN)
    double cdn_s, chn_s;              // lg15 / lu12 skin drag: Cd_from_z0(zu, 0.69e-3), Eq.11-12
    double lg15_log_ratio;            // LOG(10/z0f) / LOG(zu/z0f)
    double an05_us_c;                 // 0.035 LOG(10/8e-4) / LOG(zu/8e-4)
    int nb_iter;
};

struct IceOut {
    double Cd, Ch, Ce, t_zu, q_zu, Ub;
    double CdN, ChN, CeN, z0, us, L, UN10, CdN_frm;
    bool bad;
};

// frice: this point's ice fraction (lu12); frice_form: the fraction the LG15 form drag is computed from -- the LAST
// point's in the reference, because CdN_f_LG15_light assigns its whole result array inside the point loop
// (mod_cdn_form_ice.f90:324)
template <int IALGO, bool ZTEQ>
ABD IceOut solve_ice(const IceUniform &u, double Ts, double t_zt, double qs, double q_zt, double U_zu, double frice,
                     double frice_form)
{
    IceOut o;
    o.bad = false;
    o.CdN_frm = 0.;
    o.Ub = abm::dmax(U_zu, WSPD_THRSHLD_ICE);
    o.t_zu = abm::dmax(t_zt, 100.);
    o.q_zu = abm::dmax(q_zt, 0.1e-6);
    double dt = floor_abs(o.t_zu - Ts, 1.E-6);
    double dq = floor_abs(o.q_zu - qs, 1.E-9);

    if (IALGO == ICE_NEMO || IALGO == ICE_LU12) {
        // mod_blk_ice_nemo.f90:36-153 (constant 1.4e-3) / mod_blk_ice_lu12.f90:50-214 (skin + LU13 form drag, neutral)
        double Cd = RCD_ICE;
        if (IALGO == ICE_LU12) {
            o.CdN_frm = CdN10_f_LU13(frice);
            Cd = u.cdn_s + o.CdN_frm;
        }
        o.Cd = o.Ch = o.Ce = o.CdN = o.ChN = o.CeN = Cd;
        const double sq = abm::fast_sqrt(Cd);
        o.z0 = u.zu * abm::dexp(-fdiv(VKARMN, sq));
        o.us = sq * o.Ub;
        const double cs = fdiv(Cd, sq);
        o.L = abm::fast_rcp(one_on_L(o.t_zu, o.q_zu, o.us, cs * dt, cs * dq));
        o.UN10 = sq * o.Ub * INV_VKARMN * abm::dlog(fdiv(10., o.z0));
        return o;
    }

    if (IALGO == ICE_EASY) {
        // mod_blk_ice_easy.f90:35-209
        const double CdN = u.cxn[0], ChN = u.cxn[1], CeN = u.cxn[2];
        o.Cd = CdN; o.Ch = ChN; o.Ce = CeN;
        double us = 0., ts = 0., qst = 0., psim_u = 0.;
#pragma unroll 1
        for (int jit = 1; jit <= u.nb_iter; ++jit) {
            const double dt_zu = o.t_zu - Ts, dq_zu = o.q_zu - qs;       // no floor here (:148-149)
            const double sq = abm::fast_sqrt(o.Cd);
            us = sq * o.Ub;
            const double r = abm::fast_rcp(abm::dmax(sq, 1.E-15));
            ts = o.Ch * dt_zu * r;
            qst = o.Ce * dq_zu * r;
            const double r1oL = one_on_L(o.t_zu, o.q_zu, us, ts, qst);
            const double zeta_u = clip_abs(u.zu * r1oL, 50.0);
            psim_u = psi_m_ice(zeta_u);
            const double psih_u = psi_h_ice(zeta_u);
            double x = 1. + u.sqrt_cdn * INV_VKARMN * (u.log_zu10 - psim_u);
            o.Cd = abm::dmin(abm::dmax(fdiv(CdN, x * x), CX_MIN), 1.9E-3);
            x = fdiv((u.log_zu10 - psih_u) * INV_VKARMN, u.sqrt_cdn);
            const double y = fdiv(abm::fast_sqrt(o.Cd), u.sqrt_cdn);
            o.Ch = abm::dmin(abm::dmax(fdiv(ChN * y, 1. + ChN * x), CX_MIN), 1.9E-3);
            o.Ce = abm::dmin(abm::dmax(fdiv(CeN * y, 1. + CeN * x), CX_MIN), 1.9E-3);
            if (!ZTEQ) {
                const double zeta_t = clip_abs(u.zt * r1oL, 50.0);
                const double c = psih_u - psi_h_ice(zeta_t) + u.log_ztu;
                o.t_zu = t_zt - ts * INV_VKARMN * c;
                o.q_zu = abm::dmax(0., q_zt - qst * INV_VKARMN * c);
            }
        }
        o.CdN = CdN; o.ChN = ChN; o.CeN = CeN;
        o.z0 = u.zu * abm::dexp(-(fdiv(VKARMN, abm::fast_sqrt(o.Cd)) + psim_u));
        o.us = us;
        o.L = abm::fast_rcp(one_on_L(o.t_zu, o.q_zu, us, ts, qst));
        o.UN10 = abm::fast_sqrt(o.Cd) * o.Ub * INV_VKARMN * abm::dlog(fdiv(10., o.z0));
        return o;
    }

    if (IALGO == ICE_AN05) {
        // mod_blk_ice_an05.f90:41-243
        const double nu = visc_air(o.t_zu);
        double us = u.an05_us_c * o.Ub;
        double z0 = rough_leng_m(us, nu);
        double log_z0 = abm::dlog(z0);
#pragma unroll 1
        for (int jit = 1; jit <= 2; ++jit) {
            us = abm::dmax(fdiv(o.Ub * VKARMN, u.log_zu - log_z0), 1.E-9);
            z0 = rough_leng_m(us, nu);
            log_z0 = abm::dlog(z0);
        }
        double log_z0t, log_z0q;
        log_rough_leng_tq(z0, log_z0, us, nu, log_z0t, log_z0q, o.bad);
        double ts = fdiv(dt * VKARMN, u.log_zu - log_z0t);
        double qst = fdiv(dq * VKARMN, u.log_zu - log_z0q);
#pragma unroll 1
        for (int jit = 1; jit <= u.nb_iter; ++jit) {
            const double r1oL = one_on_L(o.t_zu, o.q_zu, us, ts, qst);
            const double zeta_u = clip_abs(u.zu * r1oL, 50.0);
            z0 = rough_leng_m(us, nu);
            log_z0 = abm::dlog(z0);
            log_rough_leng_tq(z0, log_z0, us, nu, log_z0t, log_z0q, o.bad);
            const double psih_u = psi_h_ice(zeta_u);
            ts = fdiv(dt * VKARMN, u.log_zu - log_z0t - psih_u);
            qst = fdiv(dq * VKARMN, u.log_zu - log_z0q - psih_u);
            us = abm::dmax(fdiv(o.Ub * VKARMN, u.log_zu - log_z0 - psi_m_ice(zeta_u)), 1.E-9);
            if (!ZTEQ) {
                const double zeta_t = clip_abs(u.zt * r1oL, 50.0);
                const double c = u.log_ztu + psih_u - psi_h_ice(zeta_t);
                o.t_zu = t_zt - ts * INV_VKARMN * c;
                o.q_zu = q_zt - qst * INV_VKARMN * c;
                dt = floor_abs(o.t_zu - Ts, 1.E-6);
                dq = floor_abs(o.q_zu - qs, 1.E-9);
            }
        }
        const double x = fdiv(us, o.Ub);
        o.Cd = x * x;
        o.Ch = fdiv(x * ts, dt);
        o.Ce = fdiv(x * qst, dq);
        const double r = abm::fast_rcp(u.log_zu - log_z0);
        o.CdN = VKARMN2 * r * r;
        o.ChN = fdiv(VKARMN2 * r, u.log_zu - log_z0t);
        o.CeN = fdiv(VKARMN2 * r, u.log_zu - log_z0q);
        o.z0 = z0;
        o.us = us;
        o.L = abm::fast_rcp(one_on_L(o.t_zu, o.q_zu, us, ts, qst));
        o.UN10 = us * INV_VKARMN * (u.log_zu - u.log_zu10 - log_z0);      // LOG(10/z0)
        return o;
    }

    // ICE_LG15: mod_blk_ice_lg15.f90:53-307 (and the over-ice part of mod_blk_ice_lg15_io.f90:39-370)
    {
        const double z0_s = RZ0_I_S_0, z0_f = RZ0_I_F_0;
        const double CdN_s = u.cdn_s, ChN_s = u.chn_s;
        const double CdN_f = CdN_f_LG15_light(u.lg15_log_ratio, frice_form);
        const double ChN_f = fdiv(CdN_f, 1. + LOG_5 * INV_VKARMN * abm::fast_sqrt(CdN_f));
        const double CdN = CdN_s + CdN_f, z0_tot = z0_s + z0_f;
        o.Cd = CdN;
        o.Ch = ChN_s + ChN_f;
        double RiB = ri_bulk(u.zt, Ts, t_zt, qs, q_zt, o.Ub);
#pragma unroll 1
        for (int jit = 1; jit <= u.nb_iter; ++jit) {
            double wnd_zt = o.Ub;
            if (!ZTEQ) {
                const double c = u.log_ztu + f_h_louis(u.zu, RiB, CdN, z0_tot) - f_h_louis(u.zt, RiB, CdN, z0_tot);
                wnd_zt = abm::dmin(abm::dmax(o.Ub + (abm::fast_sqrt(o.Cd) * o.Ub) * c, WSPD_THRSHLD_ICE), o.Ub);
            }
            RiB = ri_bulk(u.zt, Ts, t_zt, qs, q_zt, wnd_zt);
            o.Cd = CdN_s * f_m_louis(u.zu, RiB, CdN_s, z0_s);
            o.Ch = ChN_s * f_h_louis(u.zu, RiB, CdN_s, z0_s);
            o.Cd = o.Cd + CdN_f * f_m_louis(u.zu, RiB, CdN_f, z0_f);
            o.Ch = o.Ch + ChN_f * f_h_louis(u.zu, RiB, CdN_f, z0_f);
            if (!ZTEQ) {
                const double c = u.log_ztu + f_h_louis(u.zu, RiB, CdN, z0_tot) - f_h_louis(u.zt, RiB, CdN, z0_tot);
                const double r = abm::fast_rcp(abm::fast_sqrt(o.Cd));
                o.t_zu = t_zt - (o.Ch * dt * r) * INV_VKARMN * c;
                o.q_zu = abm::dmax(0., q_zt - (o.Ch * dq * r) * INV_VKARMN * c);
                dt = floor_abs(o.t_zu - Ts, 1.E-6);
                dq = floor_abs(o.q_zu - qs, 1.E-9);
            }
        }
        o.Ce = o.Ch;
        o.CdN_frm = CdN_f;
        o.CdN = CdN;
        o.ChN = o.CeN = ChN_s + ChN_f;
        o.z0 = u.zu * abm::dexp(-fdiv(VKARMN, abm::fast_sqrt(CdN)));
        const double sq = abm::fast_sqrt(o.Cd);
        o.us = sq * o.Ub;
        o.L = abm::fast_rcp(one_on_L(o.t_zu, o.q_zu, sq * o.Ub, fdiv(o.Ch * dt, sq), fdiv(o.Ce * dq, sq)));
        o.UN10 = sq * o.Ub * INV_VKARMN * abm::dlog(fdiv(10., o.z0));
        return o;
    }
}

}  // namespace abd

#undef ABD
