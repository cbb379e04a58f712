/*
 * aerobulk_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C restatement of the reference algorithm behind `aerobulk_model`
 * (brodeau/aerobulk).  See aerobulk_oracle.h for scope and parity status.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src/).  Operation order, literals (1.7320508, .3333, 0.6667,
 * 9.8 vs 9.80665 ...) and quirks are kept on purpose: "fixing" them breaks
 * parity with the reference.
 *
 * Conventions of the Fortran source that are mirrored here:
 *   SIGN(a,b)    -> copysign(fabs(a), b)          (so SIGN(0.5,+0.)=+0.5)
 *   x**y (real)  -> pow(x,y)   x**2 -> x*x        (gfortran lowering)
 *   MODULO       -> floored modulo                INT() -> truncation
 *   a*b/c, a/b*c -> left to right, no re-association, no FMA
 */
#include "aerobulk_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* Fortran intrinsics                                                  */
/* ------------------------------------------------------------------ */
static inline double SIGN(double a, double b) { return copysign(fabs(a), b); }
static inline double MAX(double a, double b) { return (a > b) ? a : b; }
static inline double MIN(double a, double b) { return (a < b) ? a : b; }
static inline double MODULO(double a, double p)
{
    double r = fmod(a, p);
    if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
    return r;
}
static inline int IMODULO(int a, int p)
{
    int r = a % p;
    if (r != 0 && ((r < 0) != (p < 0))) r += p;
    return r;
}

/* ------------------------------------------------------------------ */
/* mod_const.f90:38-120                                                */
/* ------------------------------------------------------------------ */
static const double grav = 9.8;                       /* :38 */
static const double rpi = 3.141592653589793;          /* :39 */
static const double roce_alb0 = 0.066;                /* :49 */
static const double emiss_w = 0.98;                   /* :55 */
static const double stefan = 5.67E-8;                 /* :57 */
static const double rt0 = 273.15;                     /* :60 */
static const double rCp0_w = 4190.;                   /* :63 */
static const double rho0_w = 1025.;                   /* :64 */
static const double rnu0_w = 1.e-6;                   /* :65 */
static const double rk0_w = 0.6;                      /* :66 */
static const double rCp_dry = 1005.0;                 /* :71 */
static const double rCp_vap = 1860.0;                 /* :72 */
static const double R_dry = 287.05;                   /* :74 */
static const double R_vap = 461.495;                  /* :75 */
static const double R_gas = 8.314510;                 /* :76 */
static const double rmm_dryair = 28.9647e-3;          /* :78 */
static const double rmm_water = 18.0153e-3;           /* :79 */
static const double rLevap = 2.46e+6;                 /* :91 */
/* Patm = 101000. (:98) is only the default of pot_temp; pPref is always given on this path */
static const double rho0_a = 1.2;                     /* :99 */
static const double vkarmn = 0.4;                     /* :103 */
static const double rdct_qsat_salt = 0.98;            /* :105 */
static const double z0_sea_max = 0.0025;              /* :106 */
static const double Cx_min = 0.1E-3;                  /* :114 */
static const double ref_tau_max = 10.;                /* :149 */

/* derived PARAMETERs: gfortran folds them operation by operation in FP64,
 * which is what evaluating the same expressions at start-up gives */
static double vkarmn2;     /* :104  0.4*0.4 */
static double rpoiss_dry;  /* :82   R_dry/rCp_dry */
static double rgamma_dry;  /* :83   grav/rCp_dry */
static double reps0;       /* :86   R_dry/R_vap */
static double rctv0;       /* :87   R_vap/R_dry - 1 */
static double rcst_cs;     /* :109 */
static double sq_radrw;    /* :112  SQRT(rho0_a/rho0_w) */
static int consts_ready = 0;

static void init_consts(void)
{
    if (consts_ready) return;
    vkarmn2 = 0.4 * 0.4;
    rpoiss_dry = R_dry / rCp_dry;
    rgamma_dry = grav / rCp_dry;
    reps0 = R_dry / R_vap;
    rctv0 = R_vap / R_dry - 1.;
    rcst_cs = -16. * 9.80665 * rho0_w * rCp0_w * rnu0_w * rnu0_w * rnu0_w / (rk0_w * rk0_w);
    sq_radrw = sqrt(rho0_a / rho0_w);
    consts_ready = 1;
}

/* sanity ranges, mod_const.f90:138-146 */
static const double ref_sst_min = 270., ref_sst_max = 320.;
static const double ref_taa_min = 180., ref_taa_max = 330.;
static const double ref_sha_min = 0., ref_sha_max = 0.08;
static const double ref_dpt_min = 150., ref_dpt_max = 330.;
static const double ref_rlh_min = 0., ref_rlh_max = 100.;
static const double ref_slp_min = 80000., ref_slp_max = 110000.;
static const double ref_wnd_min = 0., ref_wnd_max = 50.;
static const double ref_rsw_min = 0., ref_rsw_max = 1500.0;
static const double ref_rlw_min = 0., ref_rlw_max = 750.0;

/* ------------------------------------------------------------------ */
/* mod_phymbl.f90                                                      */
/* ------------------------------------------------------------------ */

/* pot_temp_sclr, mod_phymbl.f90:163-187 (pPref always given on this path) */
static double pot_temp(double pTa, double pPz, double pPref)
{
    return pTa * pow(pPref / pPz, rpoiss_dry);
}

/* virt_temp_sclr, mod_phymbl.f90:247-269 */
static double virt_temp(double pTa, double pqa) { return pTa * (1. + rctv0 * pqa); }

/* e_sat_sclr, mod_phymbl.f90:777-800 (Goff 1957; rt0 not rtt0 on purpose) */
static double e_sat(double pTa)
{
    double zta = MAX(pTa, 180.);
    double ztmp = rt0 / zta;
    return 100. * (pow(10., 10.79574 * (1. - ztmp) - 5.028 * log10(zta / rt0)
                                + 1.50475 * 1.e-4 * (1. - pow(10., -8.2969 * (zta / rt0 - 1.)))
                                + 0.42873 * 1.e-3 * (pow(10., 4.76955 * (1. - ztmp)) - 1.) + 0.78614));
}

/* q_sat_sclr, mod_phymbl.f90:881-904 (l_ice never true on the ocean path) */
static double q_sat(double pTa, double pslp)
{
    double ze_s = e_sat(pTa);
    return reps0 * ze_s / (pslp - (1. - reps0) * ze_s);
}

/* Pz_from_P0_tz_qz_sclr, mod_phymbl.f90:283-318 (3 fixed iterations) */
static double Pz_from_P0_tz_qz(double pz, double pslp, double pTa, double pqa)
{
    double zpa = pslp;
    for (int it = 1; it <= 3; it++) {
        double zqsat = q_sat(pTa, zpa);
        double zf = pqa / zqsat;
        double zxm = (1. - zf) * rmm_dryair + zf * rmm_water;
        zpa = pslp * exp(-grav * zxm * pz / (R_gas * pTa));
    }
    return zpa;
}

/* Theta_from_z_P0_T_q_sclr, mod_phymbl.f90:343-365 */
static double Theta_from_z_P0_T_q(double pz, double pslp, double pTa, double pqa)
{
    double zPz = Pz_from_P0_tz_qz(pz, pslp, pTa, pqa);
    return pot_temp(pTa, zPz, pslp);
}

/* rho_air_sclr, mod_phymbl.f90:522-537 */
static double rho_air(double pTa, double pqa, double pslp)
{
    return MAX(pslp / (R_dry * pTa * (1. + rctv0 * pqa)), 0.8);
}

/* visc_air_sclr, mod_phymbl.f90:549-563 */
static double visc_air(double pTa)
{
    double ztc = pTa - rt0;
    double ztc2 = ztc * ztc;
    return 1.326e-5 * (1. + 6.542E-3 * ztc + 8.301e-6 * ztc2 - 4.84e-9 * ztc2 * ztc);
}

/* L_vap_sclr, mod_phymbl.f90:579-592 */
static double L_vap(double psst) { return (2.501 - 0.00237 * (psst - rt0)) * 1.e6; }

/* cp_air_sclr, mod_phymbl.f90:603-616 */
static double cp_air(double pqa) { return rCp_dry + rCp_vap * pqa; }

/* gamma_moist_sclr, mod_phymbl.f90:627-649 */
static double gamma_moist(double pTa, double pqa)
{
    double zta = MAX(pTa, 180.);
    double zqa = MAX(pqa, 1.E-6);
    double zwa = zqa / (1. - zqa);
    double ziRT = 1. / (R_dry * zta);
    double zLvap = L_vap(pTa);
    return grav * (1. + zLvap * zwa * ziRT) / (rCp_dry + zLvap * zLvap * zwa * reps0 * ziRT / zta);
}

/* One_on_L_sclr, mod_phymbl.f90:666-693 */
static double One_on_L(double pThta, double pqa, double pus, double pts, double pqs)
{
    double zqa = (1. + rctv0 * pqa);
    double r = grav * vkarmn * (pts * zqa + rctv0 * pThta * pqs) / MAX(pus * pus * pThta * zqa, 1.E-9);
    return SIGN(MIN(fabs(r), 200.), r);
}

/* Ri_bulk_sclr, mod_phymbl.f90:712-747 (pTa_layer/pqa_layer never given here) */
static double Ri_bulk(double pz, double psst, double pThta, double pssq, double pqa, double pub)
{
    double zsstv = virt_temp(psst, pssq);
    double zdthv = virt_temp(pThta, pqa) - zsstv;
    double ztv = 0.5 * (zsstv + virt_temp(pThta - rgamma_dry * pz, pqa));
    return grav * zdthv * pz / (ztv * pub * pub);
}

/* q_air_rh, mod_phymbl.f90:963-985 (prha in %) */
static double q_air_rh(double prha, double pTa, double pslp)
{
    double ze = 0.01 * prha * e_sat(pTa);
    return ze * reps0 / MAX(pslp - (1. - reps0) * ze, 1.);
}

/* q_air_dp, mod_phymbl.f90:990-1000 */
static double q_air_dp(double da, double slp)
{
    double q = MAX(e_sat(da), 0.);
    return q * reps0 / MAX(slp - (1. - reps0) * q, 1.);
}

/* alpha_sw_sclr, mod_phymbl.f90:1267-1280 */
static double alpha_sw(double psst) { return 2.1e-5 * pow(MAX(psst - rt0 + 3.2, 0.), 0.79); }

/* qlw_net_sclr, mod_phymbl.f90:1291-1314 (over water) */
static double qlw_net(double pdwlw, double pts)
{
    double zt2 = pts * pts;
    return emiss_w * (pdwlw - stefan * zt2 * zt2);
}

/* BULK_FORMULA_SCLR, mod_phymbl.f90:1149-1203 (l_ice false) */
static void bulk_formula(double pzu, double pts, double pqs, double pThta, double pqa,
                         double pCd, double pCh, double pCe, double pwnd, double pUb, double pslp,
                         double *pTau, double *pQsen, double *pQlat, double *pEvap, double *prhoa)
{
    double zta = pThta - rgamma_dry * pzu;
    double zrho = rho_air(zta, pqa, pslp);
    zrho = rho_air(zta, pqa, pslp - zrho * grav * pzu);
    double zUrho = pUb * MAX(zrho, 1.);
    *pTau = zUrho * pCd * pwnd;
    double zevap = zUrho * pCe * (pqa - pqs);
    *pQsen = zUrho * pCh * (pThta - pts) * cp_air(pqa);
    *pQlat = L_vap(pts) * zevap;
    if (pEvap) *pEvap = zevap;
    if (prhoa) *prhoa = zrho;
}

/* UPDATE_QNSOL_TAU_SCLR, mod_phymbl.f90:1059-1103 */
static void update_qnsol_tau(double pzu, double pts, double pqs, double pThta, double pqa,
                             double pust, double ptst, double pqst, double pwnd, double pUb,
                             double pslp, double prlw, double *pQns, double *pTau, double *Qlat)
{
    double zdt = pThta - pts;
    zdt = SIGN(MAX(fabs(zdt), 1.E-09), zdt);
    double zdq = pqa - pqs;
    zdq = SIGN(MAX(fabs(zdq), 1.E-12), zdq);
    double zz0 = pust / pUb;
    double zCd = zz0 * zz0;
    double zCh = zz0 * ptst / zdt;
    double zCe = zz0 * pqst / zdq;
    double zQsen, zQlat;
    bulk_formula(pzu, pts, pqs, pThta, pqa, zCd, zCh, zCe, pwnd, pUb, pslp, pTau, &zQsen, &zQlat, NULL, NULL);
    double zQlw = qlw_net(prlw, pts);
    *pQns = zQlat + zQsen + zQlw;
    if (Qlat) *Qlat = zQlat;
}

/* z0_from_Cd_sclr, mod_phymbl.f90:1335-1352 */
static double z0_from_Cd_psi(double pzu, double pCd, double ppsi)
{
    return pzu * exp(-(vkarmn / sqrt(pCd) + ppsi));
}
static double z0_from_Cd_neutral(double pzu, double pCd) { return pzu * exp(-vkarmn / sqrt(pCd)); }

/* UN10_from_CD_sclr, mod_phymbl.f90:1532-1547 */
static double UN10_from_CD(double pzu, double pUb, double pCd, double ppsi)
{
    return sqrt(pCd) * pUb / vkarmn * log(10. / z0_from_Cd_psi(pzu, pCd, ppsi));
}

/* UN10_from_ustar, mod_phymbl.f90:1498-1510 */
static double UN10_from_ustar(double pzu, double pUzu, double pus, double ppsi)
{
    return pUzu - pus / vkarmn * (log(pzu / 10.) - ppsi);
}

/* z0tq_LKB, mod_phymbl.f90:1635-1701 (Liu-Katsaros-Businger table) */
static double z0tq_LKB(int iflag, double pRer, double pz0)
{
    static const double XA[2][8] = {
        {0.177, 1.376, 1.026, 1.625, 4.661, 34.904, 1667.19, 5.88e5},
        {0.292, 1.808, 1.393, 1.956, 4.994, 30.709, 1448.68, 2.98e5}};
    static const double XB[2][8] = {
        {0., 0.929, -0.599, -1.018, -1.475, -2.067, -2.907, -3.935},
        {0., 0.826, -0.528, -0.870, -1.297, -1.845, -2.682, -3.616}};
    static const double XRAN[9] = {0., 0.11, 0.825, 3.0, 10.0, 30.0, 100., 300., 1000.};
    double r = -999.;
    double zrr = pRer;
    if ((zrr > 0.) && (zrr < 1000.)) {
        int jm = 0, lfound = 0;
        while (!lfound) {
            jm = jm + 1;
            lfound = ((zrr > XRAN[jm - 1]) && (zrr <= XRAN[jm]));
        }
        r = XA[iflag - 1][jm - 1] * pow(zrr, XB[iflag - 1][jm - 1]) * pz0 / zrr;
    }
    return MIN(MAX(fabs(r), 1.E-9), 0.05);
}

/* delta_skin_layer_sclr, mod_phymbl.f90:2010-2046 */
static double delta_skin_layer(double palpha, double pQd, double pustar_a, int has_qlat, double Qlat)
{
    double zQd = pQd;
    if (has_qlat) zQd = pQd + 0.026 * MIN(Qlat, 0.) * rCp0_w / rLevap / palpha;
    double ztf = 0.5 + SIGN(0.5, zQd);
    double zusw = MAX(pustar_a, 1.E-4) * sq_radrw;
    double zusw2 = zusw * zusw;
    double zlamb = 6. * pow(1. + pow(MAX(palpha * rcst_cs / (zusw2 * zusw2) * zQd, 0.), 0.75), (-1. / 3.));
    double ztmp = rnu0_w / zusw;
    return (1. - ztf) * zlamb * ztmp + ztf * MIN(6. * ztmp, 0.007);
}

/* ------------------------------------------------------------------ */
/* mod_common_coare.f90                                                */
/* ------------------------------------------------------------------ */

/* psi_m_coare_sclr, mod_common_coare.f90:217-254 */
static double psi_m_coare(double pzeta)
{
    double zphi_m = pow(fabs(1. - 15. * pzeta), .25);
    double zpsi_k = 2. * log((1. + zphi_m) / 2.) + log((1. + zphi_m * zphi_m) / 2.) - 2. * atan(zphi_m) + 0.5 * rpi;
    double zphi_c = pow(fabs(1. - 10.15 * pzeta), .3333);
    double zpsi_c = 1.5 * log((1. + zphi_c + zphi_c * zphi_c) / 3.) - 1.7320508 * atan((1. + 2. * zphi_c) / 1.7320508) + 1.813799447;
    double zf = pzeta * pzeta;
    zf = zf / (1. + zf);
    double zc = MIN(50., 0.35 * pzeta);
    double zstb = 0.5 + SIGN(0.5, pzeta);
    return (1. - zstb) * ((1. - zf) * zpsi_k + zf * zpsi_c)
           - zstb * (1. + 1. * pzeta + 0.6667 * (pzeta - 14.28) / exp(zc) + 8.525);
}

/* psi_h_coare_sclr, mod_common_coare.f90:305-344 */
static double psi_h_coare(double pzeta)
{
    double zphi_h = pow(fabs(1. - 15. * pzeta), .5);
    double zpsi_k = 2. * log((1. + zphi_h) / 2.);
    double zphi_c = pow(fabs(1. - 34.15 * pzeta), .3333);
    double zpsi_c = 1.5 * log((1. + zphi_c + zphi_c * zphi_c) / 3.) - 1.7320508 * atan((1. + 2. * zphi_c) / 1.7320508) + 1.813799447;
    double zf = pzeta * pzeta;
    zf = zf / (1. + zf);
    double zc = MIN(50., 0.35 * pzeta);
    double zstb = 0.5 + SIGN(0.5, pzeta);
    return (1. - zstb) * ((1. - zf) * zpsi_k + zf * zpsi_c)
           - zstb * (pow(fabs(1. + 2. * pzeta / 3.), 1.5) + .6667 * (pzeta - 14.28) / exp(zc) + 8.525);
}

/* FIRST_GUESS_COARE_SCLR, mod_common_coare.f90:33-179 */
static void first_guess_coare(double zt, double zu, double psst, double t_zt, double pssq, double q_zt,
                              double U_zu, double pcharn, double *pus, double *pts, double *pqs,
                              double *t_zu, double *q_zu, double *Ubzu, double *pz0)
{
    const double zzi0 = 600., zBeta0 = 1.2;
    int l_zt_equal_zu = (fabs(zu - zt) < 0.01);

    *t_zu = MAX(t_zt, 180.);
    *q_zu = MAX(q_zt, 1.e-6);

    double zz0 = 0.0001;

    double zlog_10 = log(10.);
    double zlog_zt = log(zt);
    double zlog_zu = log(zu);
    double zc_a = 0.035 * log(10. / zz0) / log(zu / zz0);
    double zc_b = 0.004 * zzi0 * zBeta0 * zBeta0 * zBeta0;

    double zdt = *t_zu - psst;
    zdt = SIGN(MAX(fabs(zdt), 1.E-09), zdt);
    double zdq = *q_zu - pssq;
    zdq = SIGN(MAX(fabs(zdq), 1.E-12), zdq);

    double zNu_a = visc_air(*t_zu);

    double zUb = sqrt(U_zu * U_zu + 0.5 * 0.5);

    double zus = zc_a * zUb;

    zz0 = pcharn * zus * zus / grav + 0.11 * zNu_a / zus;
    zz0 = MIN(MAX(fabs(zz0), 1.E-8), 1.);
    double zlog_z0 = log(zz0);

    double zq = vkarmn / (zlog_zu - zlog_z0);
    double zCd = zq * zq;
    double z1_o_sqrt_Cd10 = (zlog_10 - zlog_z0) / vkarmn;

    double zz0t = 10. / exp(vkarmn / (0.00115 * z1_o_sqrt_Cd10));
    zz0t = MIN(MAX(fabs(zz0t), 1.E-8), 1.);
    double zlog_z0t = log(zz0t);

    double zRib = Ri_bulk(zu, psst, *t_zu, pssq, *q_zu, zUb);

    double zcc = vkarmn2 / (zCd * (zlog_zt - zlog_z0t));
    double zcc_ri = zcc * zRib;
    double z1_o_Ribcu = -zc_b / zu;
    double zstab = 0.5 + SIGN(0.5, zRib);
    double zzeta_u = (1. - zstab) * zcc_ri / (1. + zRib * z1_o_Ribcu)
                     + zstab * (zcc_ri + 27. / 9. * zRib * zRib);

    zus = MAX(zUb * vkarmn / (zlog_zu - zlog_z0 - psi_m_coare(zzeta_u)), 1.E-9);
    double ztmp = vkarmn / (zlog_zu - zlog_z0t - psi_h_coare(zzeta_u));
    double zts = zdt * ztmp;
    double zqs = zdq * ztmp;

    if (!l_zt_equal_zu) {
        double zzeta_t = zt * zzeta_u / zu;
        double zprf = log(zt / zu) + psi_h_coare(zzeta_u) - psi_h_coare(zzeta_t);
        *t_zu = t_zt - zts / vkarmn * zprf;
        *q_zu = q_zt - zqs / vkarmn * zprf;
        *q_zu = (0.5 + SIGN(0.5, *q_zu)) * *q_zu;
        zdt = *t_zu - psst;
        zdt = SIGN(MAX(fabs(zdt), 1.E-09), zdt);
        zdq = *q_zu - pssq;
        zdq = SIGN(MAX(fabs(zdq), 1.E-12), zdq);
        zts = zdt * ztmp;
        zqs = zdq * ztmp;
    }

    *pus = zus;
    *pts = zts;
    *pqs = zqs;
    *Ubzu = zUb;

    zz0 = pcharn * zus * zus / grav + 0.11 * zNu_a / zus;
    *pz0 = MIN(MAX(fabs(zz0), 1.E-8), 1.);
}

/* ------------------------------------------------------------------ */
/* persistent warm-layer state of one point                            */
/* (module arrays of mod_skin_coare.f90:31-36 / mod_skin_ecmwf.f90:52-55)*/
/* ------------------------------------------------------------------ */
typedef struct {
    double *dT_wl, *Hz_wl, *Qnt_ac, *Tau_ac;
} wl_state;

/* ------------------------------------------------------------------ */
/* mod_skin_coare.f90                                                  */
/* ------------------------------------------------------------------ */

/* CS_COARE, mod_skin_coare.f90:48-93 */
static double cs_coare(double pQsw, double pQnsol, double pustar, double pSST, double pQlat)
{
    double zQabs = pQnsol;
    double zdelta = delta_skin_layer(alpha_sw(pSST), zQabs, pustar, 1, pQlat);
    for (int jc = 1; jc <= 4; jc++) {
        double zfr = MAX(0.137 + 11. * zdelta - 6.6E-5 / zdelta * (1. - exp(-zdelta / 8.E-4)), 0.01);
        zQabs = pQnsol + zfr * pQsw;
        zdelta = delta_skin_layer(alpha_sw(pSST), zQabs, pustar, 1, pQlat);
    }
    return zQabs * zdelta / rk0_w;
}

/* WL_COARE, mod_skin_coare.f90:97-250 */
static void wl_coare(wl_state st, double pQsw, double pQnsol, double pTau, double pSST, double plon,
                     int isd, int iwait, double rdt, double gdept)
{
    const double Hwl_max = 20., Rich0 = 0.65, zfr0 = 0.5;
    double zQabs = 0.;
    double zfr = zfr0;
    int l_exit = 0, l_destroy_wl = 0;

    double zdTwl = *st.dT_wl;
    double zHwl = MAX(MIN(*st.Hz_wl, Hwl_max), 0.1);
    double zqac = *st.Qnt_ac;
    double ztac = *st.Tau_ac;

    /* local solar time, :146-150 */
    double rlag_gw_h = -1. * MODULO((360. - MODULO(plon, 360.)) / 15., 24.);
    rlag_gw_h = -1. * SIGN(MIN(fabs(rlag_gw_h), fabs(MODULO(rlag_gw_h, 24.))), rlag_gw_h + 12.);
    int ilag_gw_s = (int)(rlag_gw_h * 3600.);
    int isd_sol = IMODULO(isd + ilag_gw_s, 24 * 3600);
    double rhr_sol = (double)isd_sol / 3600.;

    double zalpha = alpha_sw(pSST);
    double zcd1 = sqrt(2. * Rich0 * rCp0_w / (zalpha * grav * rho0_w));
    double zcd2 = sqrt(2. * zalpha * grav / (Rich0 * rho0_w)) / (pow(rCp0_w, 1.5));

    if ((rhr_sol > 4.) && (rhr_sol <= 6.5)) {
        l_exit = 1;
        l_destroy_wl = 1;
    }

    if (!l_exit) {
        zfr = 1. - (0.28 * 0.014 * (1. - exp(-zHwl / 0.014)) + 0.27 * 0.357 * (1. - exp(-zHwl / 0.357))
                    + 0.45 * 12.82 * (1 - exp(-zHwl / 12.82))) / zHwl;
        zQabs = zfr * pQsw + pQnsol;
        if ((fabs(zdTwl) < 1.E-6) && (zQabs <= 0.)) l_exit = 1;
    }

    if ((!l_exit) && (*st.Qnt_ac + zQabs * rdt <= 0.)) {
        l_exit = 1;
        l_destroy_wl = 1;
    }

    if (!l_exit) {
        ztac = *st.Tau_ac + MAX(.002, pTau) * rdt;
        for (int jl = 1; jl <= 5; jl++) {
            zfr = 1. - (0.28 * 0.014 * (1. - exp(-zHwl / 0.014)) + 0.27 * 0.357 * (1. - exp(-zHwl / 0.357))
                        + 0.45 * 12.82 * (1 - exp(-zHwl / 12.82))) / zHwl;
            zQabs = zfr * pQsw + pQnsol;
            zqac = *st.Qnt_ac + zQabs * rdt;
            if (zqac <= 0.) break;
            zHwl = MAX(MIN(Hwl_max, zcd1 * ztac / sqrt(zqac)), 0.1);
        }
        if (zqac <= 0.) {
            l_destroy_wl = 1;
            l_exit = 1;
        } else {
            zdTwl = zcd2 * pow(zqac, 1.5) / ztac * MAX(zqac / fabs(zqac), 0.);
            double flg = 0.5 + SIGN(0.5, gdept - zHwl);
            zdTwl = zdTwl * (flg + (1. - flg) * gdept / zHwl);
        }
    }

    if (l_destroy_wl) {
        zdTwl = 0.;
        zfr = 0.75;
        zHwl = Hwl_max;
        zqac = 0.;
        ztac = 0.;
    }
    (void)zfr;

    if (iwait == 0) {
        *st.dT_wl = zdTwl;
        *st.Hz_wl = zHwl;
        *st.Qnt_ac = zqac;
        *st.Tau_ac = ztac;
    }
}

/* ------------------------------------------------------------------ */
/* mod_skin_ecmwf.f90                                                  */
/* ------------------------------------------------------------------ */

/* CS_ECMWF, mod_skin_ecmwf.f90:68-110 */
static double cs_ecmwf(double pQsw, double pQnsol, double pustar, double pSST)
{
    double zQabs = pQnsol;
    double zdelta = delta_skin_layer(alpha_sw(pSST), zQabs, pustar, 0, 0.);
    for (int jc = 1; jc <= 4; jc++) {
        double zfr = MAX(0.065 + 11. * zdelta - 6.6E-5 / zdelta * (1. - exp(-zdelta / 8.E-4)), 0.01);
        zQabs = pQnsol + zfr * pQsw;
        zdelta = delta_skin_layer(alpha_sw(pSST), zQabs, pustar, 0, 0.);
    }
    return zQabs * zdelta / rk0_w;
}

/* PHI, mod_skin_ecmwf.f90:233-253 (Takaya et al. 2010 Eq.5) */
static double PHI(double pzeta)
{
    double zzt2 = pzeta * pzeta;
    double ztf = 0.5 + SIGN(0.5, pzeta);
    return ztf * (1. + (5. * pzeta + 4. * zzt2) / (1. + 3. * pzeta + 0.25 * zzt2))
           + (1. - ztf) * 1. / sqrt(1. - 16. * (-fabs(pzeta)));
}

/* WL_ECMWF, mod_skin_ecmwf.f90:113-230 (pustk never given on this path) */
static void wl_ecmwf(wl_state st, double pQsw, double pQnsol, double pustar, double pSST, double rdt, double gdept)
{
    const double rNuwl0 = 0.5;
    const double zRhoCp_w = rho0_w * rCp0_w;

    double zHwl = *st.Hz_wl;
    double flg = 0.5 + SIGN(0.5, gdept - zHwl);
    double ztcorr = flg + (1. - flg) * gdept / zHwl;
    double zdTwl_b = MAX(*st.dT_wl / ztcorr, 0.);

    double zalpha = alpha_sw(pSST);

    double zfr = 1. - 0.28 * exp(-71.5 * zHwl) - 0.27 * exp(-2.8 * zHwl) - 0.45 * exp(-0.07 * zHwl);
    double zQabs = zfr * pQsw + pQnsol;

    double zusw = MAX(pustar, 1.E-4) * sq_radrw;
    double zusw2 = zusw * zusw;

    double zla = 0.3;
    double zfLa = MAX(pow(zla, (-2. / 3.)), 1.);

    double zwf = 0.5 + SIGN(0.5, zQabs);

    double zcst1 = vkarmn * grav * zalpha;
    double zL2 = zcst1 * zQabs / (zRhoCp_w * zusw2 * zusw);
    double zcst2 = zcst1 / (5. * zHwl * zusw2);
    double zcst0 = rdt * (rNuwl0 + 1.) / zHwl;
    double zA = zcst0 * zQabs / (rNuwl0 * zRhoCp_w);
    double zcst3 = -zcst0 * vkarmn * zusw * zfLa;

    double zdTwl_n = zdTwl_b;
    for (int jc = 1; jc <= 10; jc++) {
        zdTwl_n = 0.5 * (zdTwl_n + zdTwl_b);
        double zL1 = sqrt(zdTwl_n * zcst2);
        double zeta = (1. - zwf) * zHwl * zL1 + zwf * zHwl * zL2;
        double zB = zcst3 / PHI(zeta);
        zdTwl_n = MAX(zdTwl_b + zA + zB * zdTwl_n, 0.);
    }
    *st.dT_wl = zdTwl_n * ztcorr;
}

/* ------------------------------------------------------------------ */
/* per-point interface shared by the five TURB_* restatements           */
/* ------------------------------------------------------------------ */
typedef struct {
    /* in/out */
    double T_s, q_s;
    /* out */
    double Cd, Ch, Ce, t_zu, q_zu, Ubzu;
    /* optional out */
    double CdN, ChN, CeN, z0, us, L, UN10, dT_cs;
} turb_io;

typedef struct {
    int l_use_cs, l_use_wl;
    double Qsw, rad_lw, slp, plong;
    int isd;
    double rdt, gdept;
    wl_state st;
} skin_in;

/* ------------------------------------------------------------------ */
/* mod_blk_ncar.f90                                                    */
/* ------------------------------------------------------------------ */

/* cd_n10_ncar_sclr, mod_blk_ncar.f90:244-271 */
static double cd_n10_ncar(double pw10)
{
    double zw = pw10;
    double zw6 = zw * zw * zw;
    zw6 = zw6 * zw6;
    double zgt33 = 0.5 + SIGN(0.5, (zw - 33.));
    double r = 1.e-3 * ((1. - zgt33) * (2.7 / zw + 0.142 + zw / 13.09 - 3.14807E-10 * zw6) + zgt33 * 2.34);
    return MAX(r, Cx_min);
}
/* ch_n10_ncar_sclr :287-302, ce_n10_ncar_sclr :313-322 */
static double ch_n10_ncar(double psqrtcdn10, double pstab)
{
    return MAX(1.e-3 * psqrtcdn10 * (18. * pstab + 32.7 * (1. - pstab)), Cx_min);
}
static double ce_n10_ncar(double psqrtcdn10) { return MAX(1.e-3 * (34.6 * psqrtcdn10), Cx_min); }

/* psi_m_ncar_sclr, mod_blk_ncar.f90:333-363 */
static double psi_m_ncar(double pzeta)
{
    double zta = pzeta;
    double zx2 = sqrt(fabs(1. - 16. * zta));
    zx2 = MAX(zx2, 1.);
    double zx = sqrt(zx2);
    double zpsi_unst = 2. * log((1. + zx) * 0.5) + log((1. + zx2) * 0.5) - 2. * atan(zx) + rpi * 0.5;
    double zpsi_stab = -5. * zta;
    double zstab = 0.5 + SIGN(0.5, zta);
    return zstab * zpsi_stab + (1. - zstab) * zpsi_unst;
}
/* psi_h_ncar_sclr, mod_blk_ncar.f90:379-407 */
static double psi_h_ncar(double pzeta)
{
    double zta = pzeta;
    double zx2 = sqrt(fabs(1. - 16. * zta));
    zx2 = MAX(zx2, 1.);
    double zpsi_unst = 2. * log(0.5 * (1. + zx2));
    double zpsi_stab = -5. * zta;
    double zstab = 0.5 + SIGN(0.5, zta);
    return zstab * zpsi_stab + (1. - zstab) * zpsi_unst;
}

/* turb_ncar, mod_blk_ncar.f90:57-240 (one point) */
static void turb_ncar(int nb_iter, double zt, double zu, double sst, double t_zt, double ssq, double q_zt,
                      double U_zu, turb_io *o)
{
    int l_zt_equal_zu = (fabs(zu - zt) < 0.01);
    double Ubzu = MAX(0.5, U_zu);
    double zlog1 = log(zt / zu);
    double zlog2 = log(zu / 10.);

    double zstab = 0.5 + SIGN(0.5, virt_temp(t_zt, q_zt) - virt_temp(sst, ssq));
    double zCdN = cd_n10_ncar(Ubzu);
    double zsqrt_CdN = sqrt(zCdN);
    double Cd = zCdN;
    double Ce = ce_n10_ncar(zsqrt_CdN);
    double Ch = ch_n10_ncar(zsqrt_CdN, zstab);
    double zsqrt_Cd = zsqrt_CdN;
    double t_zu = MAX(t_zt, 180.);
    double q_zu = MAX(q_zt, 1.e-6);
    double zus = 0., z1oL = 0., zUn10 = 0., zChN = 0., zCeN = 0.;

    for (int jit = 1; jit <= nb_iter; jit++) {
        double zdt = t_zu - sst;
        double zdq = q_zu - ssq;
        zus = zsqrt_Cd * Ubzu;
        double zts = Ch / zsqrt_Cd * zdt;
        double zqs = Ce / zsqrt_Cd * zdq;
        z1oL = One_on_L(t_zu, q_zu, zus, zts, zqs);
        double zeta_u = zu * z1oL;
        zeta_u = SIGN(MIN(fabs(zeta_u), 10.), zeta_u);
        if (!l_zt_equal_zu) {
            double zeta_t = zt * z1oL;
            zeta_t = SIGN(MIN(fabs(zeta_t), 10.), zeta_t);
            double ztmp = zlog1 + psi_h_ncar(zeta_u) - psi_h_ncar(zeta_t);
            t_zu = t_zt - zts / vkarmn * ztmp;
            q_zu = q_zt - zqs / vkarmn * ztmp;
            q_zu = MAX(0., q_zu);
        }
        double zpsi_m = psi_m_ncar(zeta_u);
        zUn10 = MAX(0.25, UN10_from_CD(zu, Ubzu, Cd, zpsi_m));
        zCdN = cd_n10_ncar(zUn10);
        zsqrt_CdN = sqrt(zCdN);
        double ztmp = 1. + zsqrt_CdN / vkarmn * (zlog2 - zpsi_m);
        Cd = MAX(zCdN / (ztmp * ztmp), Cx_min);
        zsqrt_Cd = sqrt(Cd);
        ztmp = (zlog2 - psi_h_ncar(zeta_u)) / vkarmn / zsqrt_CdN;
        double ztmp2 = zsqrt_Cd / zsqrt_CdN;
        zstab = 0.5 + SIGN(0.5, zeta_u);
        zChN = 1.e-3 * zsqrt_CdN * (18. * zstab + 32.7 * (1. - zstab));
        zCeN = 1.e-3 * (34.6 * zsqrt_CdN);
        Ch = MAX(zChN * ztmp2 / (1. + zChN * ztmp), Cx_min);
        Ce = MAX(zCeN * ztmp2 / (1. + zCeN * ztmp), Cx_min);
    }
    o->Cd = Cd; o->Ch = Ch; o->Ce = Ce; o->t_zu = t_zu; o->q_zu = q_zu; o->Ubzu = Ubzu;
    o->CdN = zCdN; o->CeN = zCeN; o->ChN = zChN; o->UN10 = zUn10; o->L = 1. / z1oL; o->us = zus;
    o->z0 = MIN(z0_from_Cd_neutral(zu, zCdN), z0_sea_max);
    o->dT_cs = 0.;
}

/* ------------------------------------------------------------------ */
/* mod_blk_coare3p0.f90 / mod_blk_coare3p6.f90                          */
/* ------------------------------------------------------------------ */

/* charn_coare3p0, mod_blk_coare3p0.f90:420-447 */
static double charn_coare3p0(double pwnd)
{
    double zw = pwnd;
    double zgt10 = 0.5 + SIGN(0.5, (zw - 10.));
    double zgt18 = 0.5 + SIGN(0.5, (zw - 18.));
    return (1. - zgt10) * 0.011
           + zgt10 * ((1. - zgt18) * (0.011 + (0.018 - 0.011) * (zw - 10.) / (18. - 10.)) + zgt18 * (0.018));
}
/* charn_coare3p6_sclr, mod_blk_coare3p6.f90:417-432 */
static double charn_coare3p6(double pwnd) { return MAX(MIN(0.0017 * pwnd - 0.005, 0.028), 0.); }

/* Test-only knob.  doc/ex_ab.dat predates the switch of mod_blk_coare3p0.f90:237 from
 * visc_air(t_zu) to visc_air(theta_zt); with the knob on, the oracle reproduces the
 * file's COARE 3.0 rows to every printed digit, which pins everything else of that path. */
static int dbg_coare3p0_visc_tzu = 0;
void abo_debug_coare3p0_visc_at_tzu(int on) { dbg_coare3p0_visc_tzu = on; }

/* turb_coare3p0, mod_blk_coare3p0.f90:54-358 (one point; array prologue :207-214 folded in) */
static void turb_coare3p0(int nb_iter, double zt, double zu, double t_zt, double q_zt, double U_zu,
                          const skin_in *sk, turb_io *o)
{
    const double zi0 = 600., Beta0 = 1.25, zeta_abs_max = 50.;
    double zm_ztzu = (fabs(zu - zt) < 0.01) ? 0. : 1.;
    int l_use_cs = sk->l_use_cs, l_use_wl = sk->l_use_wl;
    int l_skin = l_use_cs || l_use_wl;

    double zSST = o->T_s;
    double zT_s = o->T_s, zq_s = o->q_s;
    if (l_skin) {
        if (l_use_cs) zT_s = zT_s - 0.25;
        zq_s = rdct_qsat_salt * q_sat(MAX(zT_s, 200.), sk->slp);
    }

    double zlog_10 = log(10.);
    double zlog_zt = log(zt);
    double zlog_zu = log(zu);

    double zt_zt = t_zt, zq_zt = q_zt, zUzu = U_zu;
    double zus, zts, zqs, zt_zu, zq_zu, zUbzu, zz0;
    first_guess_coare(zt, zu, zT_s, zt_zt, zq_s, zq_zt, zUzu, charn_coare3p0(zUzu),
                      &zus, &zts, &zqs, &zt_zu, &zq_zu, &zUbzu, &zz0);

    double zlog_z0 = log(zz0);
    double znu_a = visc_air(zt_zt); /* :237 -- theta at zt, unlike COARE 3.6 */
    if (dbg_coare3p0_visc_tzu) znu_a = visc_air(zt_zu); /* test-only: the variant doc/ex_ab.dat was captured with */

    double zdt = zt_zu - zT_s;
    zdt = SIGN(MAX(fabs(zdt), 1.E-09), zdt);
    double zdq = zq_zu - zq_s;
    zdq = SIGN(MAX(fabs(zdq), 1.E-12), zdq);

    double z1oL = 0., zlog_z0t = 0., zdT_cs = 0.;
    for (int jit = 1; jit <= nb_iter; jit++) {
        double zus2 = zus * zus;
        z1oL = One_on_L(zt_zu, zq_zu, zus, zts, zqs);
        z1oL = SIGN(MIN(fabs(z1oL), 200.), z1oL);

        double zgust2 = Beta0 * Beta0 * zus2 * pow(MAX(-zi0 * z1oL / vkarmn, 0.), (2. / 3.));
        zUbzu = MAX(sqrt(zUzu * zUzu + zgust2), 0.2);

        double zzta_u = zu * z1oL;
        zzta_u = SIGN(MIN(fabs(zzta_u), zeta_abs_max), zzta_u);
        double zzta_t = zt * z1oL;
        zzta_t = SIGN(MIN(fabs(zzta_t), zeta_abs_max), zzta_t);

        double zUn10 = zus / vkarmn * (zlog_10 - zlog_z0);
        zz0 = charn_coare3p0(zUn10) * zus2 / grav + 0.11 * znu_a / zus;
        zz0 = MIN(MAX(fabs(zz0), 1.E-9), 1.);
        zlog_z0 = log(zz0);

        double ztmp1 = pow(znu_a / (zz0 * zus), 0.6);
        double zz0t = MIN(1.1E-4, 5.5E-5 * ztmp1);
        zz0t = MIN(MAX(fabs(zz0t), 1.E-9), 1.);
        zlog_z0t = log(zz0t);

        double ztmp0 = psi_h_coare(zzta_u);
        ztmp1 = vkarmn / (zlog_zu - zlog_z0t - ztmp0);
        zts = zdt * ztmp1;
        zqs = zdq * ztmp1;
        zus = MAX(zUbzu * vkarmn / (zlog_zu - zlog_z0 - psi_m_coare(zzta_u)), 1.E-9);

        ztmp1 = zlog_zt - zlog_zu + ztmp0 - psi_h_coare(zzta_t);
        zt_zu = zt_zt - zm_ztzu * zts / vkarmn * ztmp1;
        zq_zu = zq_zt - zm_ztzu * zqs / vkarmn * ztmp1;

        if (l_use_cs) {
            double zQns, zTau, zQlat;
            update_qnsol_tau(zu, zT_s, zq_s, zt_zu, zq_zu, zus, zts, zqs, zUzu, zUbzu, sk->slp, sk->rad_lw,
                             &zQns, &zTau, &zQlat);
            zdT_cs = cs_coare(sk->Qsw, zQns, zus, zSST, zQlat);
            zT_s = zSST + zdT_cs;
            if (l_use_wl) zT_s = zT_s + *sk->st.dT_wl;
            zq_s = rdct_qsat_salt * q_sat(MAX(zT_s, 200.), sk->slp);
        }
        if (l_use_wl) {
            double zQns, zTau;
            update_qnsol_tau(zu, zT_s, zq_s, zt_zu, zq_zu, zus, zts, zqs, zUzu, zUbzu, sk->slp, sk->rad_lw,
                             &zQns, &zTau, NULL);
            wl_coare(sk->st, sk->Qsw, zQns, zTau, zSST, sk->plong, sk->isd, nb_iter % jit, sk->rdt, sk->gdept);
            zT_s = zSST + *sk->st.dT_wl;
            if (l_use_cs) zT_s = zT_s + zdT_cs;
            zq_s = rdct_qsat_salt * q_sat(MAX(zT_s, 200.), sk->slp);
        }
        zdt = zt_zu - zT_s;
        zdt = SIGN(MAX(fabs(zdt), 1.E-09), zdt);
        zdq = zq_zu - zq_s;
        zdq = SIGN(MAX(fabs(zdq), 1.E-12), zdq);
    }

    o->T_s = zT_s; o->q_s = zq_s; o->t_zu = zt_zu; o->q_zu = zq_zu; o->Ubzu = zUbzu;
    double ztmp0 = zus / zUbzu;
    o->Cd = MAX(ztmp0 * ztmp0, Cx_min);
    o->Ch = MAX(ztmp0 * zts / zdt, Cx_min);
    o->Ce = MAX(ztmp0 * zqs / zdq, Cx_min);
    ztmp0 = 1. / (zlog_zu - zlog_z0);
    o->CdN = MAX(vkarmn2 * ztmp0 * ztmp0, Cx_min);
    double ztmp1 = vkarmn2 * ztmp0 / (zlog_zu - zlog_z0t);
    o->ChN = MAX(ztmp1, Cx_min);
    o->CeN = MAX(ztmp1, Cx_min);
    o->z0 = zz0; o->us = zus; o->L = 1. / z1oL; o->UN10 = zus / vkarmn * (zlog_10 - zlog_z0);
    o->dT_cs = zdT_cs;
}

/* TURB_COARE3P6, mod_blk_coare3p6.f90:123-413 (one point; array prologue :271-276 folded in) */
static void turb_coare3p6(int nb_iter, double zt, double zu, double t_zt, double q_zt, double U_zu,
                          const skin_in *sk, turb_io *o)
{
    const double zi0 = 600., Beta0 = 1.2, zeta_abs_max = 50.;
    int l_zt_equal_zu = (fabs(zu - zt) < 0.01);
    int l_use_cs = sk->l_use_cs, l_use_wl = sk->l_use_wl;

    double xSST = o->T_s;
    double T_s = o->T_s, q_s = o->q_s;
    if (l_use_cs || l_use_wl) {
        if (l_use_cs) T_s = T_s - 0.25;
        q_s = rdct_qsat_salt * q_sat(MAX(T_s, 200.), sk->slp);
    }

    double zlog_10 = log(10.);
    double zlog_zt = log(zt);
    double zlog_zu = log(zu);

    double zUzu = U_zu;
    double zus, zts, zqs, t_zu, q_zu, Ubzu, zz0;
    first_guess_coare(zt, zu, T_s, t_zt, q_s, q_zt, zUzu, charn_coare3p6(zUzu),
                      &zus, &zts, &zqs, &t_zu, &q_zu, &Ubzu, &zz0);

    double zlog_z0 = log(zz0);
    double znu_a = visc_air(t_zu); /* :294 */

    double zdt = t_zu - T_s;
    zdt = SIGN(MAX(fabs(zdt), 1.E-09), zdt);
    double zdq = q_zu - q_s;
    zdq = SIGN(MAX(fabs(zdq), 1.E-12), zdq);

    double z1oL = 0., zlog_z0t = 0., zdT_cs = 0., zzta_t = 0.;
    for (int jit = 1; jit <= nb_iter; jit++) {
        double zus2 = zus * zus;
        z1oL = One_on_L(t_zu, q_zu, zus, zts, zqs);
        z1oL = SIGN(MIN(fabs(z1oL), 200.), z1oL);

        double zgust2 = Beta0 * Beta0 * zus2 * pow(MAX(-zi0 * z1oL / vkarmn, 0.), (2. / 3.));
        Ubzu = MAX(sqrt(zUzu * zUzu + zgust2), 0.2);

        double zzta_u = zu * z1oL;
        zzta_u = SIGN(MIN(fabs(zzta_u), zeta_abs_max), zzta_u);
        if (!l_zt_equal_zu) {
            zzta_t = zt * z1oL;
            zzta_t = SIGN(MIN(fabs(zzta_t), zeta_abs_max), zzta_t);
        }

        double zUn10 = zus / vkarmn * (zlog_10 - zlog_z0);
        zz0 = charn_coare3p6(zUn10) * zus2 / grav + 0.11 * znu_a / zus;
        zz0 = MIN(MAX(fabs(zz0), 1.E-9), 1.);
        zlog_z0 = log(zz0);

        double ztmp1 = pow(znu_a / (zz0 * zus), 0.72);
        double zz0t = MIN(1.6E-4, 5.8E-5 * ztmp1);
        zz0t = MIN(MAX(fabs(zz0t), 1.E-9), 1.);
        zlog_z0t = log(zz0t);

        double ztmp0 = psi_h_coare(zzta_u);
        ztmp1 = vkarmn / (zlog_zu - zlog_z0t - ztmp0);
        zts = zdt * ztmp1;
        zqs = zdq * ztmp1;
        zus = MAX(Ubzu * vkarmn / (zlog_zu - zlog_z0 - psi_m_coare(zzta_u)), 1.E-9);

        if (!l_zt_equal_zu) {
            ztmp1 = zlog_zt - zlog_zu + ztmp0 - psi_h_coare(zzta_t);
            t_zu = t_zt - zts / vkarmn * ztmp1;
            q_zu = q_zt - zqs / vkarmn * ztmp1;
        }

        if (l_use_cs) {
            double zQns, zTau, zQlat;
            update_qnsol_tau(zu, T_s, q_s, t_zu, q_zu, zus, zts, zqs, zUzu, Ubzu, sk->slp, sk->rad_lw,
                             &zQns, &zTau, &zQlat);
            zdT_cs = cs_coare(sk->Qsw, zQns, zus, xSST, zQlat);
            T_s = xSST + zdT_cs;
            if (l_use_wl) T_s = T_s + *sk->st.dT_wl;
            q_s = rdct_qsat_salt * q_sat(MAX(T_s, 200.), sk->slp);
        }
        if (l_use_wl) {
            double zQns, zTau;
            update_qnsol_tau(zu, T_s, q_s, t_zu, q_zu, zus, zts, zqs, zUzu, Ubzu, sk->slp, sk->rad_lw,
                             &zQns, &zTau, NULL);
            wl_coare(sk->st, sk->Qsw, zQns, zTau, xSST, sk->plong, sk->isd, nb_iter % jit, sk->rdt, sk->gdept);
            T_s = xSST + *sk->st.dT_wl;
            if (l_use_cs) T_s = T_s + zdT_cs;
            q_s = rdct_qsat_salt * q_sat(MAX(T_s, 200.), sk->slp);
        }
        if (l_use_cs || l_use_wl || (!l_zt_equal_zu)) {
            zdt = t_zu - T_s;
            zdt = SIGN(MAX(fabs(zdt), 1.E-09), zdt);
            zdq = q_zu - q_s;
            zdq = SIGN(MAX(fabs(zdq), 1.E-12), zdq);
        }
    }

    o->T_s = T_s; o->q_s = q_s; o->t_zu = t_zu; o->q_zu = q_zu; o->Ubzu = Ubzu;
    double ztmp0 = zus / Ubzu;
    o->Cd = MAX(ztmp0 * ztmp0, Cx_min);
    o->Ch = MAX(ztmp0 * zts / zdt, Cx_min);
    o->Ce = MAX(ztmp0 * zqs / zdq, Cx_min);
    ztmp0 = 1. / (zlog_zu - zlog_z0);
    o->CdN = MAX(vkarmn2 * ztmp0 * ztmp0, Cx_min);
    double ztmp1 = vkarmn2 * ztmp0 / (zlog_zu - zlog_z0t);
    o->ChN = MAX(ztmp1, Cx_min);
    o->CeN = MAX(ztmp1, Cx_min);
    o->z0 = zz0; o->us = zus; o->L = 1. / z1oL; o->UN10 = zus / vkarmn * (zlog_10 - zlog_z0);
    o->dT_cs = zdT_cs;
}

/* ------------------------------------------------------------------ */
/* mod_blk_ecmwf.f90                                                   */
/* ------------------------------------------------------------------ */

/* cap_zeta, mod_blk_ecmwf.f90:551-564 */
static double cap_zeta(double pzeta)
{
    double zta = MAX(pzeta, -50.);
    zta = MIN(zta, 5.);
    return zta;
}

/* psi_m_ecmwf_scl, mod_blk_ecmwf.f90:441-477 */
static double psi_m_ecmwf(double pzeta)
{
    double zc = 5. / 0.35;
    double zta = cap_zeta(pzeta);
    double zx2 = sqrt(fabs(1. - 16. * zta));
    double zx = sqrt(zx2);
    double ztmp = 1. + zx;
    double zpsi_unst = log(0.125 * ztmp * ztmp * (1. + zx2)) - 2. * atan(zx) + 0.5 * rpi;
    double zpsi_stab = -(2. / 3. * (zta - zc) * exp(-0.35 * zta)) - zta - 2. / 3. * zc;
    double zstab = 0.5 + SIGN(0.5, zta);
    return zstab * zpsi_stab + (1. - zstab) * zpsi_unst;
}

/* psi_h_ecmwf_scl, mod_blk_ecmwf.f90:498-533 */
static double psi_h_ecmwf(double pzeta)
{
    double zc = 5. / 0.35;
    double zta = cap_zeta(pzeta);
    double zx2 = sqrt(fabs(1. - 16. * zta));
    double zpsi_unst = 2. * log(0.5 * (1. + zx2));
    double zpsi_stab = -(2. / 3. * (zta - zc) * exp(-0.35 * zta)) - pow(fabs(1. + 2. / 3. * zta), 1.5) - 2. / 3. * zc + 1.;
    double zstab = 0.5 + SIGN(0.5, zta);
    return zstab * zpsi_stab + (1. - zstab) * zpsi_unst;
}

/* turb_ecmwf, mod_blk_ecmwf.f90:63-383 (one point; array prologue :209-216 folded in) */
static void turb_ecmwf(int nb_iter, double zt, double zu, double t_zt, double q_zt, double U_zu,
                       const skin_in *sk, turb_io *o)
{
    const double charn0_ecmwf = 0.018, zi0 = 1000., Beta0 = 1., alpha_M = 0.11, alpha_H = 0.40, alpha_Q = 0.62;
    double zm_ztzu = (fabs(zu - zt) < 0.01) ? 0. : 1.;
    int l_use_cs = sk->l_use_cs, l_use_wl = sk->l_use_wl;
    int l_skin = l_use_cs || l_use_wl;

    double zSST = o->T_s;
    double zT_s = o->T_s, zq_s = o->q_s;
    if (l_skin) {
        if (l_use_cs) zT_s = zT_s - 0.25;
        zq_s = rdct_qsat_salt * q_sat(MAX(zT_s, 200.), sk->slp);
    }

    double zlog_10 = log(10.);
    double zlog_zu = log(zu);
    double zlog_ztu = log(zt / zu);

    double zt_zt = t_zt, zq_zt = q_zt, zUzu = U_zu;
    double zus, zts, zqs, zt_zu, zq_zu, zUbzu, zz0;
    first_guess_coare(zt, zu, zT_s, zt_zt, zq_s, zq_zt, zUzu, charn0_ecmwf,
                      &zus, &zts, &zqs, &zt_zu, &zq_zu, &zUbzu, &zz0);

    double zlog_z0 = log(zz0);
    double znu_a = visc_air(zt_zt);

    double zdt = zt_zu - zT_s;
    zdt = SIGN(MAX(fabs(zdt), 1.E-09), zdt);
    double zdq = zq_zu - zq_s;
    zdq = SIGN(MAX(fabs(zdq), 1.E-12), zdq);

    double z1oL = One_on_L(zt_zu, zq_zu, zus, zts, zqs);
    double zzeta_u = zu * z1oL;
    double zzeta_t = zt * z1oL;

    double zz0t = MIN(MAX(fabs(1. / (0.1 * exp(vkarmn / (0.00115 / (vkarmn / (zlog_10 - zlog_z0)))))), 1.E-9), 1.);
    double zlog_z0t = log(zz0t);

    double zFm = zlog_zu - zlog_z0 - psi_m_ecmwf(zzeta_u) + psi_m_ecmwf(zz0 * z1oL);
    double zpsi_h_u = psi_h_ecmwf(zzeta_u);
    double zFh = zlog_zu - zlog_z0t - zpsi_h_u + psi_h_ecmwf(zz0t * z1oL);

    double zlog_z0q = 0., zpsi_h_z0q = 0., zdT_cs = 0.;
    for (int jit = 1; jit <= nb_iter; jit++) {
        double zRib = Ri_bulk(zu, zT_s, zt_zu, zq_s, zq_zu, zUbzu);
        z1oL = zRib * zFm * zFm / zFh / zu;
        z1oL = SIGN(MIN(fabs(z1oL), 200.), z1oL);

        zzeta_u = zu * z1oL;
        double zpsi_m_u = psi_m_ecmwf(zzeta_u);
        zpsi_h_u = psi_h_ecmwf(zzeta_u);
        zzeta_t = zt * z1oL;
        double zpsi_h_t = psi_h_ecmwf(zzeta_t);

        zFm = zlog_zu - zlog_z0 - zpsi_m_u + psi_m_ecmwf(zz0 * z1oL);

        zus = zUbzu * vkarmn / zFm;
        double zus2 = zus * zus;
        double ztmp0 = znu_a / zus;
        zz0 = MIN(fabs(alpha_M * ztmp0 + charn0_ecmwf * zus2 / grav), 0.001);
        zz0t = MIN(fabs(alpha_H * ztmp0), 0.001);
        double zz0q = MIN(fabs(alpha_Q * ztmp0), 0.001);

        zlog_z0 = log(zz0);
        zlog_z0t = log(zz0t);
        zlog_z0q = log(zz0q);

        double zpsi_m_z0 = psi_m_ecmwf(zz0 * z1oL);
        double zpsi_h_z0t = psi_h_ecmwf(zz0t * z1oL);
        zpsi_h_z0q = psi_h_ecmwf(zz0q * z1oL);

        ztmp0 = Beta0 * Beta0 * zus2 * pow(MAX(-zi0 * z1oL / vkarmn, 0.), (2. / 3.));
        zUbzu = MAX(sqrt(zUzu * zUzu + ztmp0), 0.2);

        ztmp0 = zpsi_h_u - zpsi_h_z0t;
        double ztmp1 = vkarmn / (zlog_zu - zlog_z0t - ztmp0);
        zts = zdt * ztmp1;
        ztmp1 = zlog_ztu + ztmp0 - zpsi_h_t + zpsi_h_z0t;
        zt_zu = zt_zt - zm_ztzu * zts / vkarmn * ztmp1;

        ztmp0 = zpsi_h_u - zpsi_h_z0q;
        ztmp1 = vkarmn / (zlog_zu - zlog_z0q - ztmp0);
        zqs = zdq * ztmp1;
        ztmp1 = zlog_ztu + ztmp0 - zpsi_h_t + zpsi_h_z0q;
        zq_zu = MAX(zq_zt - zm_ztzu * zqs / vkarmn * ztmp1, 0.);

        zFm = zlog_zu - zlog_z0 - zpsi_m_u + zpsi_m_z0;
        zFh = zlog_zu - zlog_z0t - zpsi_h_u + zpsi_h_z0t;

        if (l_use_cs) {
            double zQns, zTau;
            update_qnsol_tau(zu, zT_s, zq_s, zt_zu, zq_zu, zus, zts, zqs, zUzu, zUbzu, sk->slp, sk->rad_lw,
                             &zQns, &zTau, NULL);
            zdT_cs = cs_ecmwf(sk->Qsw, zQns, zus, zSST);
            zT_s = zSST + zdT_cs;
            if (l_use_wl) zT_s = zT_s + *sk->st.dT_wl;
            zq_s = rdct_qsat_salt * q_sat(MAX(zT_s, 200.), sk->slp);
        }
        if (l_use_wl) {
            double zQns, zTau;
            update_qnsol_tau(zu, zT_s, zq_s, zt_zu, zq_zu, zus, zts, zqs, zUzu, zUbzu, sk->slp, sk->rad_lw,
                             &zQns, &zTau, NULL);
            wl_ecmwf(sk->st, sk->Qsw, zQns, zus, zSST, sk->rdt, sk->gdept);
            zT_s = zSST + *sk->st.dT_wl;
            if (l_use_cs) zT_s = zT_s + zdT_cs;
            zq_s = rdct_qsat_salt * q_sat(MAX(zT_s, 200.), sk->slp);
        }
        zdt = zt_zu - zT_s;
        zdt = SIGN(MAX(fabs(zdt), 1.E-09), zdt);
        zdq = zq_zu - zq_s;
        zdq = SIGN(MAX(fabs(zdq), 1.E-12), zdq);
    }

    o->T_s = zT_s; o->q_s = zq_s; o->t_zu = zt_zu; o->q_zu = zq_zu; o->Ubzu = zUbzu;
    double zFq = zlog_zu - zlog_z0q - zpsi_h_u + zpsi_h_z0q;
    o->Cd = MAX(vkarmn2 / (zFm * zFm), Cx_min);
    o->Ch = MAX(vkarmn2 / (zFm * zFh), Cx_min);
    o->Ce = MAX(vkarmn2 / (zFm * zFq), Cx_min);
    double ztmp0 = 1. / (zlog_zu - zlog_z0);
    o->CdN = MAX(vkarmn2 * ztmp0 * ztmp0, Cx_min);
    double ztmp1 = vkarmn2 * ztmp0 / (zlog_zu - zlog_z0t);
    o->ChN = MAX(ztmp1, Cx_min);
    o->CeN = MAX(ztmp1, Cx_min);
    o->z0 = zz0; o->us = zus; o->L = 1. / z1oL; o->UN10 = zus / vkarmn * (zlog_10 - zlog_z0);
    o->dT_cs = zdT_cs;
}

/* ------------------------------------------------------------------ */
/* mod_blk_andreas.f90                                                 */
/* ------------------------------------------------------------------ */

/* u_star_andreas_sclr, mod_blk_andreas.f90:275-293 */
static double u_star_andreas(double pun10)
{
    double za = pun10 - 8.271;
    double zt = za + sqrt(0.12 * za * za + 0.181);
    return 0.239 + 0.0433 * zt;
}

/* psi_m_andreas, mod_blk_andreas.f90:307-360 */
static double psi_m_andreas(double pzeta)
{
    const double zam = 5.;
    const double zbm = zam / 6.5;
    const double z1o3 = 1. / 3.;
    const double zsr3 = sqrt(3.);
    double zta = MIN(pzeta, 15.);
    double zx2 = sqrt(fabs(1. - 16. * zta));
    zx2 = MAX(zx2, 1.);
    double zx = sqrt(zx2);
    double zpsi_unst = 2. * log(fabs((1. + zx) * 0.5)) + log(fabs((1. + zx2) * 0.5)) - 2. * atan(zx) + rpi * 0.5;
    zx = pow(fabs(1. + zta), z1o3);
    double zbbm = pow(fabs((1. - zbm) / zbm), z1o3);
    double zpsi_stab = -(3. * zam / zbm * (zx - 1.)) + zam * zbbm / (2. * zbm) * (
                            2. * log(fabs((zx + zbbm) / (1. + zbbm)))
                            - log(fabs((zx * zx - zx * zbbm + zbbm * zbbm) / (1. - zbbm + zbbm * zbbm)))
                            + 2. * zsr3 * (atan((2. * zx - zbbm) / (zsr3 * zbbm)) - atan((2. - zbbm) / (zsr3 * zbbm))));
    double zstab = 0.5 + SIGN(0.5, zta);
    return zstab * zpsi_stab + (1. - zstab) * zpsi_unst;
}

/* psi_h_andreas, mod_blk_andreas.f90:363-410 */
static double psi_h_andreas(double pzeta)
{
    const double zah = 5., zbh = 5., zch = 3.;
    const double zbbh = sqrt(5.);
    double zta = MIN(pzeta, 15.);
    double zx2 = sqrt(fabs(1. - 16. * zta));
    zx2 = MAX(zx2, 1.);
    double zpsi_unst = 2. * log(0.5 * (1. + zx2));
    double zz = 2. * zta + zch;
    double zpsi_stab = -(0.5 * zbh * log(fabs(1. + zch * zta + zta * zta)))
                       + (-zah / zbbh + 0.5 * zbh * zch / zbbh)
                             * (log(fabs((zz - zbbh) / (zz + zbbh))) - log(fabs((zch - zbbh) / (zch + zbbh))));
    double zstab = 0.5 + SIGN(0.5, zta);
    return zstab * zpsi_stab + (1. - zstab) * zpsi_unst;
}

/* turb_andreas, mod_blk_andreas.f90:66-272 (whole-array statements applied to one point) */
static void turb_andreas(int nb_iter, double zt, double zu, double psst, double pt_zt, double pssq,
                         double pq_zt, double pU_zu, turb_io *o)
{
    const double rRi_max = 0.15, rCs_min = 0.35E-3;
    int l_zt_equal_zu = (fabs(zu - zt) < 0.01);

    double pUbzu = MAX(0.25, pU_zu);
    double UN10 = pUbzu;
    double pCd = 1.1E-3, pCh = 1.1E-3, pCe = 1.1E-3;
    double pt_zu = pt_zt, pq_zu = pq_zt;

    double ztmp0 = sqrt(pCd);
    double t_star = pCh / ztmp0 * (pt_zu - psst);
    double q_star = pCe / ztmp0 * (pq_zu - pssq);

    double RiB = Ri_bulk(zu, psst, pt_zu, pssq, pq_zu, pUbzu);
    double u_star = 0., zeta_u = 0., z0 = 0., ztmp1, ztmp2;

    for (int jit = 1; jit <= nb_iter; jit++) {
        if (RiB < rRi_max) u_star = u_star_andreas(UN10);
        else u_star = sqrt(Cx_min) * pUbzu;

        zeta_u = zu * One_on_L(pt_zu, pq_zu, u_star, t_star, q_star);

        ztmp0 = u_star / pUbzu;
        pCd = MAX(ztmp0 * ztmp0, Cx_min);

        z0 = MIN(z0_from_Cd_psi(zu, pCd, psi_m_andreas(zeta_u)), z0_sea_max);

        ztmp0 = z0 * u_star / visc_air(pt_zu);
        ztmp1 = z0tq_LKB(1, ztmp0, z0);
        ztmp2 = z0tq_LKB(2, ztmp0, z0);

        ztmp0 = psi_h_andreas(zeta_u);
        t_star = (pt_zu - psst) * vkarmn / (log(zu) - log(ztmp1) - ztmp0);
        q_star = (pq_zu - pssq) * vkarmn / (log(zu) - log(ztmp2) - ztmp0);

        if ((!l_zt_equal_zu) && (jit > 1)) {
            ztmp0 = zeta_u / zu * zt;
            ztmp0 = log(zt / zu) + psi_h_andreas(zeta_u) - psi_h_andreas(ztmp0);
            pt_zu = pt_zt - t_star / vkarmn * ztmp0;
            pq_zu = pq_zt - q_star / vkarmn * ztmp0;
            RiB = Ri_bulk(zu, psst, pt_zu, pssq, pq_zu, pUbzu);
        }
        UN10 = MAX(0.1, UN10_from_ustar(zu, pUbzu, u_star, psi_m_andreas(zeta_u)));
    }

    ztmp0 = u_star / pUbzu;
    pCd = MAX(ztmp0 * ztmp0, Cx_min);
    ztmp1 = pt_zu - psst;
    ztmp1 = SIGN(MAX(fabs(ztmp1), 1.E-6), ztmp1);
    ztmp2 = pq_zu - pssq;
    ztmp2 = SIGN(MAX(fabs(ztmp2), 1.E-9), ztmp2);
    pCh = MAX(ztmp0 * t_star / ztmp1, rCs_min);
    pCe = MAX(ztmp0 * q_star / ztmp2, rCs_min);

    o->Cd = pCd; o->Ch = pCh; o->Ce = pCe; o->t_zu = pt_zu; o->q_zu = pq_zu; o->Ubzu = pUbzu;
    ztmp0 = 1. / log(zu / z0);
    o->CdN = MAX(vkarmn2 * ztmp0 * ztmp0, Cx_min);
    ztmp1 = z0 * u_star / visc_air(pt_zu);
    o->ChN = vkarmn2 * ztmp0 / log(zu / z0tq_LKB(1, ztmp1, z0));
    o->CeN = vkarmn2 * ztmp0 / log(zu / z0tq_LKB(2, ztmp1, z0));
    o->z0 = z0; o->us = u_star; o->L = zu / zeta_u;
    o->UN10 = UN10_from_ustar(zu, pUbzu, u_star, psi_m_andreas(zeta_u));
    o->dT_cs = 0.;
}

/* ------------------------------------------------------------------ */
/* session = the SAVEd module state of mod_const / mod_skin_*           */
/* ------------------------------------------------------------------ */
struct abo_session {
    int nb_iter;             /* mod_const.f90:33 */
    int nitend;              /* :22 */
    int l_use_skin_schemes;  /* :24 */
    char ctype_humidity[3];  /* :27 */
    double rdt, gdept;       /* :31-32 */
    /* warm-layer module arrays; coare and ecmwf own separate copies in the
     * reference (mod_skin_coare.f90:31-36, mod_skin_ecmwf.f90:52-55) */
    long n_coare, n_ecmwf;
    double *c_dT_wl, *c_Hz_wl, *c_Qnt_ac, *c_Tau_ac;
    double *e_dT_wl, *e_Hz_wl;
    int nthreads;
    char errmsg[512];
};

abo_session *abo_new(void)
{
    init_consts();
    abo_session *s = (abo_session *)calloc(1, sizeof(*s));
    s->nb_iter = 5;
    s->nitend = 1;
    s->l_use_skin_schemes = 0;
    strcpy(s->ctype_humidity, "sh");
    s->rdt = 3600.;
    s->gdept = 1.;
    s->nthreads = 1;
    return s;
}

static void free_coare_state(abo_session *s)
{
    free(s->c_dT_wl); free(s->c_Hz_wl); free(s->c_Qnt_ac); free(s->c_Tau_ac);
    s->c_dT_wl = s->c_Hz_wl = s->c_Qnt_ac = s->c_Tau_ac = NULL;
    s->n_coare = 0;
}
static void free_ecmwf_state(abo_session *s)
{
    free(s->e_dT_wl); free(s->e_Hz_wl);
    s->e_dT_wl = s->e_Hz_wl = NULL;
    s->n_ecmwf = 0;
}

void abo_free(abo_session *s)
{
    if (!s) return;
    free_coare_state(s);
    free_ecmwf_state(s);
    free(s);
}

void abo_set_rdt(abo_session *s, double rdt) { s->rdt = rdt; }
void abo_set_gdept(abo_session *s, double g) { s->gdept = g; }
void abo_set_nb_iter(abo_session *s, int n) { s->nb_iter = n; }
int abo_get_nb_iter(const abo_session *s) { return s->nb_iter; }
int abo_get_use_skin(const abo_session *s) { return s->l_use_skin_schemes; }
const char *abo_get_humidity_type(const abo_session *s) { return s->ctype_humidity; }
void abo_set_threads(abo_session *s, int n) { s->nthreads = n < 1 ? 1 : n; }
const char *abo_errmsg(const abo_session *s) { return s->errmsg; }

long abo_get_state(const abo_session *s, int which, double *out)
{
    const double *src = NULL;
    long n = 0;
    if (s->n_coare) {
        n = s->n_coare;
        src = which == 0 ? s->c_dT_wl : which == 1 ? s->c_Hz_wl : which == 2 ? s->c_Qnt_ac : which == 3 ? s->c_Tau_ac : NULL;
    } else if (s->n_ecmwf) {
        n = s->n_ecmwf;
        src = which == 0 ? s->e_dT_wl : which == 1 ? s->e_Hz_wl : NULL;
    }
    if (!src) return 0;
    memcpy(out, src, (size_t)n * sizeof(double));
    return n;
}

/* ------------------------------------------------------------------ */
/* check_unit_consistency, mod_phymbl.f90:1851-1954                     */
/* ------------------------------------------------------------------ */
static int check_unit_consistency(abo_session *s, const char *cfield, const double *X, const double *X2,
                                  const signed char *mask, long n)
{
    double zmin, zmax;
    const char *cunit;
    if (!strcmp(cfield, "sst")) { zmax = ref_sst_max; zmin = ref_sst_min; cunit = "K"; }
    else if (!strcmp(cfield, "t_air")) { zmax = ref_taa_max; zmin = ref_taa_min; cunit = "K"; }
    else if (!strcmp(cfield, "sh")) { zmax = ref_sha_max; zmin = ref_sha_min; cunit = "kg/kg"; }
    else if (!strcmp(cfield, "rh")) { zmax = ref_rlh_max; zmin = ref_rlh_min; cunit = "kg/kg"; }
    else if (!strcmp(cfield, "dp")) { zmax = ref_dpt_max; zmin = ref_dpt_min; cunit = "kg/kg"; }
    else if (!strcmp(cfield, "slp")) { zmax = ref_slp_max; zmin = ref_slp_min; cunit = "Pa"; }
    else if (!strcmp(cfield, "u10") || !strcmp(cfield, "v10")) { zmax = ref_wnd_max; zmin = -ref_wnd_max; cunit = "m/s"; }
    else if (!strcmp(cfield, "wnd")) { zmax = ref_wnd_max; zmin = ref_wnd_min; cunit = "m/s"; }
    else if (!strcmp(cfield, "rad_sw")) { zmax = ref_rsw_max; zmin = ref_rsw_min; cunit = "W/m^2"; }
    else if (!strcmp(cfield, "rad_lw")) { zmax = ref_rlw_max; zmin = ref_rlw_min; cunit = "W/m^2"; }
    else return ABO_ERR_UNITS;

    double sum = 0., cnt = 0., mx = -HUGE_VAL, mn = HUGE_VAL, amx = -HUGE_VAL, amn = HUGE_VAL;
    for (long i = 0; i < n; i++) {
        /* 'wnd' is checked on SQRT(pU*pU + pV*pV), mod_aerobulk.f90:148 */
        double v = X2 ? sqrt(X[i] * X[i] + X2[i] * X2[i]) : X[i];
        sum += v * (double)mask[i];
        cnt += (double)mask[i];
        if (mask[i]) { if (v > mx) mx = v; if (v < mn) mn = v; }
        if (v > amx) amx = v;
        if (v < amn) amn = v;
    }
    double zmean = sum / cnt;
    int bad = (mx > zmax) || (mn < zmin) || (zmean < zmin) || (zmean > zmax);
    if (bad) {
        snprintf(s->errmsg, sizeof(s->errmsg),
                 " *** ERROR (check_unit_consistency@mod_phymbl): field `%s` does not seem to be in %s !"
                 " min value = %10.3e max value = %10.3e mean value = %10.3e", cfield, cunit, amn, amx, zmean);
        return ABO_ERR_UNITS;
    }
    return ABO_OK;
}

/* type_of_humidity, mod_phymbl.f90:1957-2007 */
static int type_of_humidity(abo_session *s, const double *X, const signed char *mask, long n, char out[3])
{
    double sum = 0., cnt = 0., zmax = -HUGE_VAL, zmin = HUGE_VAL;
    for (long i = 0; i < n; i++) {
        sum += X[i] * (double)mask[i];
        cnt += (double)mask[i];
        if (mask[i] == 1) { if (X[i] > zmax) zmax = X[i]; if (X[i] < zmin) zmin = X[i]; }
    }
    double zmean = sum / cnt;
    if ((zmean >= ref_sha_min) && (zmean < ref_sha_max) && (zmin >= ref_sha_min) && (zmax < ref_sha_max)) strcpy(out, "sh");
    else if ((zmean >= ref_dpt_min) && (zmean < ref_dpt_max) && (zmin >= ref_dpt_min) && (zmax < ref_dpt_max)) strcpy(out, "dp");
    else if ((zmean >= ref_rlh_min) && (zmean <= ref_rlh_max) && (zmin >= ref_rlh_min) && (zmax <= ref_rlh_max)) strcpy(out, "rh");
    else {
        snprintf(s->errmsg, sizeof(s->errmsg),
                 "ERROR: type_of_humidity()@mod_aerobulk_compute => un-identified humidity type! mean = %g min = %g max = %g",
                 zmean, zmin, zmax);
        return ABO_ERR_HUMIDITY;
    }
    return ABO_OK;
}

/* AEROBULK_INIT, mod_aerobulk.f90:24-160 */
static int aerobulk_init(abo_session *s, int Nt, const char *calgo, const double *psst, const double *pta,
                         const double *pha, const double *pU, const double *pV, const double *pslp,
                         int lskin, const double *prsw, const double *prlw, long n)
{
    int lsrad = (prsw != NULL) && (prlw != NULL);
    if (lskin) {
        if (!((strncmp(calgo, "coar", 4) == 0) || (strcmp(calgo, "ecmwf") == 0))) {
            snprintf(s->errmsg, sizeof(s->errmsg),
                     " AEROBULK_INIT => Only `COARE*` and `ECMWF` algorithms support cool-skin & warm/layer schemes");
            return ABO_ERR_SKIN_ALGO;
        }
        if (!lsrad) {
            snprintf(s->errmsg, sizeof(s->errmsg),
                     " AEROBULK_INIT => provide SW and LW rad. input if you want to use skin schemes");
            return ABO_ERR_SKIN_NORAD;
        }
        s->l_use_skin_schemes = 1; /* :74 -- never reset */
    }
    s->nitend = Nt; /* :99 */

    signed char *imask = (signed char *)malloc((size_t)n);
    long np = 0;
    for (long i = 0; i < n; i++) {
        signed char m = 1;
        if ((psst[i] < ref_sst_min) || (psst[i] > ref_sst_max)) m = 0;
        if ((pta[i] < ref_taa_min) || (pta[i] > ref_taa_max)) m = 0;
        if ((pslp[i] < ref_slp_min) || (pslp[i] > ref_slp_max)) m = 0;
        if (sqrt(pU[i] * pU[i] + pV[i] * pV[i]) > ref_wnd_max) m = 0;
        if (lsrad) {
            if ((prsw[i] < ref_rsw_min) || (prsw[i] > ref_rsw_max)) m = 0;
            if ((prlw[i] < ref_rlw_min) || (prlw[i] > ref_rlw_max)) m = 0;
        }
        imask[i] = m;
        np += m;
    }
    int rc = ABO_OK;
    if (np <= 0) {
        snprintf(s->errmsg, sizeof(s->errmsg), "the whole domain is masked! check unit consistency of input fields");
        rc = ABO_ERR_ALL_MASKED;
    }
    if (!rc) rc = type_of_humidity(s, pha, imask, n, s->ctype_humidity); /* :127 */
    if (!rc) rc = check_unit_consistency(s, "sst", psst, NULL, imask, n);
    if (!rc) rc = check_unit_consistency(s, "t_air", pta, NULL, imask, n);
    if (!rc) rc = check_unit_consistency(s, "slp", pslp, NULL, imask, n);
    if (!rc) rc = check_unit_consistency(s, "u10", pU, NULL, imask, n);
    if (!rc) rc = check_unit_consistency(s, "v10", pV, NULL, imask, n);
    if (!rc) rc = check_unit_consistency(s, "wnd", pU, pV, imask, n);
    if (!rc) rc = check_unit_consistency(s, s->ctype_humidity, pha, NULL, imask, n);
    if (!rc && lsrad) {
        rc = check_unit_consistency(s, "rad_sw", prsw, NULL, imask, n);
        if (!rc) rc = check_unit_consistency(s, "rad_lw", prlw, NULL, imask, n);
    }
    free(imask);
    return rc;
}

static int algo_id(const char *calgo)
{
    if (!strcmp(calgo, "coare3p0")) return ABO_COARE3P0;
    if (!strcmp(calgo, "coare3p6")) return ABO_COARE3P6;
    if (!strcmp(calgo, "ncar")) return ABO_NCAR;
    if (!strcmp(calgo, "ecmwf")) return ABO_ECMWF;
    if (!strcmp(calgo, "andreas")) return ABO_ANDREAS;
    return 0;
}

/* aerobulk_compute, mod_aerobulk_compute.f90:22-213 (all steps are point-wise) */
static int aerobulk_compute(abo_session *s, int jt, const char *calgo, double zt, double zu, long n,
                            const double *sst, const double *t_zt, const double *hum_zt, const double *U_zu,
                            const double *V_zu, const double *slp, double *QL, double *QH, double *Tau_x,
                            double *Tau_y, const double *rad_sw, const double *rad_lw, double *T_s, double *Evp)
{
    int ialgo = algo_id(calgo);
    if (!ialgo) {
        snprintf(s->errmsg, sizeof(s->errmsg), "ERROR: mod_aerobulk_compute.f90 => bulk algorithm %s is unknown!!!", calgo);
        return ABO_ERR_ALGO;
    }
    int ihum = !strcmp(s->ctype_humidity, "sh") ? 0 : !strcmp(s->ctype_humidity, "dp") ? 1 : 2;
    int use_skin = s->l_use_skin_schemes && (ialgo == ABO_COARE3P0 || ialgo == ABO_COARE3P6 || ialgo == ABO_ECMWF);
    if (use_skin && !(rad_sw && rad_lw)) {
        /* the reference would dereference absent optionals here (undefined behaviour) */
        snprintf(s->errmsg, sizeof(s->errmsg), "skin schemes active (sticky l_use_skin_schemes) but rad_sw/rad_lw absent");
        return ABO_ERR_SKIN_NORAD;
    }

    /* kt==nit000 -> *_INIT: allocate + initialise warm-layer state
     * (mod_blk_coare3p6.f90:250,68-95; mod_blk_coare3p0.f90:185; mod_blk_ecmwf.f90:189,387-412) */
    if (jt == 1 && use_skin) {
        if (ialgo == ABO_ECMWF) {
            if (s->n_ecmwf) { snprintf(s->errmsg, sizeof(s->errmsg), " ECMWF_INIT => allocation of dT_wl & Hz_wl failed!"); return ABO_ERR_STATE; }
            s->n_ecmwf = n;
            s->e_dT_wl = (double *)malloc((size_t)n * sizeof(double));
            s->e_Hz_wl = (double *)malloc((size_t)n * sizeof(double));
            for (long i = 0; i < n; i++) { s->e_dT_wl[i] = 0.; s->e_Hz_wl[i] = 3.; }
        } else {
            if (s->n_coare) { snprintf(s->errmsg, sizeof(s->errmsg), " COARE_INIT => allocation of Tau_ac, Qnt_ac, dT_wl & Hz_wl failed!"); return ABO_ERR_STATE; }
            s->n_coare = n;
            s->c_dT_wl = (double *)malloc((size_t)n * sizeof(double));
            s->c_Hz_wl = (double *)malloc((size_t)n * sizeof(double));
            s->c_Qnt_ac = (double *)malloc((size_t)n * sizeof(double));
            s->c_Tau_ac = (double *)malloc((size_t)n * sizeof(double));
            for (long i = 0; i < n; i++) { s->c_Tau_ac[i] = 0.; s->c_Qnt_ac[i] = 0.; s->c_dT_wl[i] = 0.; s->c_Hz_wl[i] = 20.; }
        }
    }
    if (use_skin) {
        long have = (ialgo == ABO_ECMWF) ? s->n_ecmwf : s->n_coare;
        if (have != n) { snprintf(s->errmsg, sizeof(s->errmsg), "warm-layer state missing or of wrong size (jt=%d)", jt); return ABO_ERR_STATE; }
    }

    const int nb_iter = s->nb_iter;
    const double rdt = s->rdt, gdept = s->gdept;
    long first_bad = -1;
    double bad_tau = 0.;

#ifdef _OPENMP
#pragma omp parallel for num_threads(s->nthreads) schedule(static)
#endif
    for (long i = 0; i < n; i++) {
        /* :99-108 humidity -> specific */
        double zQzt;
        if (ihum == 0) zQzt = hum_zt[i];
        else if (ihum == 1) zQzt = q_air_dp(hum_zt[i], MAX(slp[i], 50000.));
        else zQzt = q_air_rh(hum_zt[i], t_zt[i], MAX(slp[i], 50000.));
        /* :111 */
        double zWzu = sqrt(U_zu[i] * U_zu[i] + V_zu[i] * V_zu[i]);
        /* :114 */
        double zSSQ = rdct_qsat_salt * q_sat(sst[i], slp[i]);
        /* :118 */
        double zThtzt = Theta_from_z_P0_T_q(zt, slp[i], t_zt[i], zQzt);

        turb_io o;
        memset(&o, 0, sizeof(o));
        o.T_s = sst[i];
        o.q_s = zSSQ;
        skin_in sk;
        memset(&sk, 0, sizeof(sk));
        sk.slp = slp[i];
        sk.rdt = rdt;
        sk.gdept = gdept;
        if (use_skin) {
            sk.l_use_cs = 1;
            sk.l_use_wl = 1;
            sk.Qsw = (1. - roce_alb0) * rad_sw[i]; /* :135,146,161 */
            sk.rad_lw = rad_lw[i];
            sk.isd = 12;   /* :136,146 -- seconds, hard-wired */
            sk.plong = 0.; /* :126 */
            if (ialgo == ABO_ECMWF) { sk.st.dT_wl = &s->e_dT_wl[i]; sk.st.Hz_wl = &s->e_Hz_wl[i]; }
            else { sk.st.dT_wl = &s->c_dT_wl[i]; sk.st.Hz_wl = &s->c_Hz_wl[i]; sk.st.Qnt_ac = &s->c_Qnt_ac[i]; sk.st.Tau_ac = &s->c_Tau_ac[i]; }
        }
        switch (ialgo) {
        case ABO_COARE3P0: turb_coare3p0(nb_iter, zt, zu, zThtzt, zQzt, zWzu, &sk, &o); break;
        case ABO_COARE3P6: turb_coare3p6(nb_iter, zt, zu, zThtzt, zQzt, zWzu, &sk, &o); break;
        case ABO_NCAR: turb_ncar(nb_iter, zt, zu, o.T_s, zThtzt, o.q_s, zQzt, zWzu, &o); break;
        case ABO_ECMWF: turb_ecmwf(nb_iter, zt, zu, zThtzt, zQzt, zWzu, &sk, &o); break;
        default: turb_andreas(nb_iter, zt, zu, o.T_s, zThtzt, o.q_s, zQzt, zWzu, &o); break;
        }

        /* :184-185 BULK_FORMULA_VCTR */
        double zTaum, zQH, zQL, zEvap;
        bulk_formula(zu, o.T_s, o.q_s, o.t_zu, o.q_zu, o.Cd, o.Ch, o.Ce, zWzu, o.Ubzu, slp[i],
                     &zTaum, &zQH, &zQL, &zEvap, NULL);
        QH[i] = zQH;
        QL[i] = zQL;
        if (zTaum > ref_tau_max) { /* mod_phymbl.f90:1250-1253 */
#ifdef _OPENMP
#pragma omp critical
#endif
            { if (first_bad < 0 || i < first_bad) { first_bad = i; bad_tau = zTaum; } }
        }
        /* :189-194 */
        double tx = 0., ty = 0.;
        if (zWzu > 1.E-3) {
            tx = zTaum / zWzu * U_zu[i];
            ty = zTaum / zWzu * V_zu[i];
        }
        Tau_x[i] = tx;
        Tau_y[i] = ty;
        if (T_s) T_s[i] = o.T_s; /* :206 */
        if (Evp) Evp[i] = zEvap; /* :208 */
    }

    /* kt==nitend -> *_EXIT (mod_blk_coare3p6.f90:411 etc.) */
    if (use_skin && jt == s->nitend) {
        if (ialgo == ABO_ECMWF) free_ecmwf_state(s);
        else free_coare_state(s);
    }

    if (first_bad >= 0) {
        snprintf(s->errmsg, sizeof(s->errmsg),
                 "BULK_FORMULA_VCTR()@mod_phymbl: wind stress too strong! => %8.2f N/m^2 ! At linear index %ld", bad_tau, first_bad);
        return ABO_ERR_TAU;
    }
    return ABO_OK;
}

/* AEROBULK_MODEL, mod_aerobulk.f90:176-269 */
int abo_model(abo_session *s, int jt, int Nt, const char *calgo, double zt, double zu, int Ni, int Nj,
              const double *sst, const double *t_zt, const double *hum_zt, const double *U_zu,
              const double *V_zu, const double *slp, double *QL, double *QH, double *Tau_x, double *Tau_y,
              double *Evap, const int *Niter, const int *l_use_skin, const double *rad_sw,
              const double *rad_lw, double *T_s)
{
    long n = (long)Ni * (long)Nj;
    s->errmsg[0] = 0;
    if (Niter) s->nb_iter = *Niter; /* :236 sticky */
    int lskin = l_use_skin ? (*l_use_skin != 0) : 0;
    int lsrad = (rad_sw != NULL) && (rad_lw != NULL);
    if (jt < 1) {
        snprintf(s->errmsg, sizeof(s->errmsg), "AEROBULK_MODEL => jt < 1 !?? we are in a Fortran world here...");
        return ABO_ERR_JT;
    }
    int rc;
    if (lsrad) {
        if (jt == 1) {
            /* :248 -- prsw=rad_lw: the reference passes rad_lw for BOTH radiation checks */
            rc = aerobulk_init(s, Nt, calgo, sst, t_zt, hum_zt, U_zu, V_zu, slp, lskin, rad_lw, rad_lw, n);
            if (rc) return rc;
        }
        rc = aerobulk_compute(s, jt, calgo, zt, zu, n, sst, t_zt, hum_zt, U_zu, V_zu, slp, QL, QH, Tau_x, Tau_y,
                              rad_sw, rad_lw, T_s, Evap);
    } else {
        if (jt == 1) {
            rc = aerobulk_init(s, Nt, calgo, sst, t_zt, hum_zt, U_zu, V_zu, slp, lskin, NULL, NULL, n);
            if (rc) return rc;
        }
        rc = aerobulk_compute(s, jt, calgo, zt, zu, n, sst, t_zt, hum_zt, U_zu, V_zu, slp, QL, QH, Tau_x, Tau_y,
                              NULL, NULL, NULL, Evap);
    }
    return rc;
}


/* ------------------------------------------------------------------ */
/* Direct TURB_* call on arrays: what GCMs (NEMO sbcblk) and the reference's own   */
/* test programs call -- real isecday_utc / plong, l_use_cs and l_use_wl independent, optional outputs.   */
/* Interfaces: mod_blk_coare3p6.f90:123-127, mod_blk_coare3p0.f90:54-59, mod_blk_ecmwf.f90:63-68,          */
/* mod_blk_ncar.f90:57-59, mod_blk_andreas.f90:66-68.  nitend is the mod_const global (abo_set_nitend).    */
/* opt[10] = CdN ChN CeN xz0 xu_star xL xUN10 pdT_cs pdT_wl pHz_wl, each NULL when not wanted.              */
/* ------------------------------------------------------------------ */
void abo_set_nitend(abo_session *s, int nitend) { s->nitend = nitend; }

int abo_turb(abo_session *s, const char *calgo, int kt, double zt, double zu, long n,
             double *T_s, const double *t_zt, double *q_s, const double *q_zt, const double *U_zu,
             int l_use_cs, int l_use_wl,
             double *Cd, double *Ch, double *Ce, double *t_zu, double *q_zu, double *Ubzu,
             const double *Qsw, const double *rad_lw, const double *slp, int isecday_utc, const double *plong,
             double *const *opt)
{
    init_consts();
    s->errmsg[0] = 0;
    int ialgo = algo_id(calgo);
    if (!ialgo) { snprintf(s->errmsg, sizeof(s->errmsg), "unknown algorithm %s", calgo); return ABO_ERR_ALGO; }
    int skin_algo = (ialgo == ABO_COARE3P0 || ialgo == ABO_COARE3P6 || ialgo == ABO_ECMWF);
    if (!skin_algo) { l_use_cs = 0; l_use_wl = 0; }
    if ((l_use_cs || l_use_wl) && !(Qsw && rad_lw && slp)) {
        snprintf(s->errmsg, sizeof(s->errmsg), "you need to provide Qsw, rad_lw & slp to use cool-skin / warm-layer param!");
        return ABO_ERR_SKIN_NORAD;
    }
    if (l_use_wl && ialgo != ABO_ECMWF && !plong) {
        snprintf(s->errmsg, sizeof(s->errmsg), "you need to provide Qsw, rad_lw, slp, isecday_utc & plong to use warm-layer param!");
        return ABO_ERR_SKIN_NORAD;
    }
    if (kt == 1 && l_use_wl) {   /* *_INIT */
        if (ialgo == ABO_ECMWF) {
            if (s->n_ecmwf) { snprintf(s->errmsg, sizeof(s->errmsg), "ECMWF_INIT => allocation failed"); return ABO_ERR_STATE; }
            s->n_ecmwf = n;
            s->e_dT_wl = (double *)malloc((size_t)n * sizeof(double));
            s->e_Hz_wl = (double *)malloc((size_t)n * sizeof(double));
            for (long i = 0; i < n; i++) { s->e_dT_wl[i] = 0.; s->e_Hz_wl[i] = 3.; }
        } else {
            if (s->n_coare) { snprintf(s->errmsg, sizeof(s->errmsg), "COARE_INIT => allocation failed"); return ABO_ERR_STATE; }
            s->n_coare = n;
            s->c_dT_wl = (double *)malloc((size_t)n * sizeof(double));
            s->c_Hz_wl = (double *)malloc((size_t)n * sizeof(double));
            s->c_Qnt_ac = (double *)malloc((size_t)n * sizeof(double));
            s->c_Tau_ac = (double *)malloc((size_t)n * sizeof(double));
            for (long i = 0; i < n; i++) { s->c_Tau_ac[i] = 0.; s->c_Qnt_ac[i] = 0.; s->c_dT_wl[i] = 0.; s->c_Hz_wl[i] = 20.; }
        }
    }
    if (l_use_wl) {
        long have = (ialgo == ABO_ECMWF) ? s->n_ecmwf : s->n_coare;
        if (have != n) { snprintf(s->errmsg, sizeof(s->errmsg), "warm-layer state missing (kt=%d)", kt); return ABO_ERR_STATE; }
    }
    const int nb_iter = s->nb_iter;
#ifdef _OPENMP
#pragma omp parallel for num_threads(s->nthreads) schedule(static)
#endif
    for (long i = 0; i < n; i++) {
        turb_io o;
        memset(&o, 0, sizeof(o));
        o.T_s = T_s[i];
        o.q_s = q_s[i];
        skin_in sk;
        memset(&sk, 0, sizeof(sk));
        sk.l_use_cs = l_use_cs;
        sk.l_use_wl = l_use_wl;
        sk.rdt = s->rdt;
        sk.gdept = s->gdept;
        if (l_use_cs || l_use_wl) {
            sk.Qsw = Qsw[i];
            sk.rad_lw = rad_lw[i];
            sk.slp = slp[i];
            sk.isd = isecday_utc;
            sk.plong = plong ? plong[i] : 0.;
        }
        if (l_use_wl) {
            if (ialgo == ABO_ECMWF) { sk.st.dT_wl = &s->e_dT_wl[i]; sk.st.Hz_wl = &s->e_Hz_wl[i]; }
            else { sk.st.dT_wl = &s->c_dT_wl[i]; sk.st.Hz_wl = &s->c_Hz_wl[i]; sk.st.Qnt_ac = &s->c_Qnt_ac[i]; sk.st.Tau_ac = &s->c_Tau_ac[i]; }
        }
        switch (ialgo) {
        case ABO_COARE3P0: turb_coare3p0(nb_iter, zt, zu, t_zt[i], q_zt[i], U_zu[i], &sk, &o); break;
        case ABO_COARE3P6: turb_coare3p6(nb_iter, zt, zu, t_zt[i], q_zt[i], U_zu[i], &sk, &o); break;
        case ABO_NCAR: turb_ncar(nb_iter, zt, zu, o.T_s, t_zt[i], o.q_s, q_zt[i], U_zu[i], &o); break;
        case ABO_ECMWF: turb_ecmwf(nb_iter, zt, zu, t_zt[i], q_zt[i], U_zu[i], &sk, &o); break;
        default: turb_andreas(nb_iter, zt, zu, o.T_s, t_zt[i], o.q_s, q_zt[i], U_zu[i], &o); break;
        }
        T_s[i] = o.T_s; q_s[i] = o.q_s;
        Cd[i] = o.Cd; Ch[i] = o.Ch; Ce[i] = o.Ce; t_zu[i] = o.t_zu; q_zu[i] = o.q_zu; Ubzu[i] = o.Ubzu;
        if (opt) {
            if (opt[0]) opt[0][i] = o.CdN;
            if (opt[1]) opt[1][i] = o.ChN;
            if (opt[2]) opt[2][i] = o.CeN;
            if (opt[3]) opt[3][i] = o.z0;
            if (opt[4]) opt[4][i] = o.us;
            if (opt[5]) opt[5][i] = o.L;
            if (opt[6]) opt[6][i] = o.UN10;
            if (opt[7] && l_use_cs) opt[7][i] = o.dT_cs;
            if (opt[8] && l_use_wl) opt[8][i] = *sk.st.dT_wl;
            if (opt[9] && l_use_wl) opt[9][i] = *sk.st.Hz_wl;
        }
    }
    if (l_use_wl && kt == s->nitend) {   /* *_EXIT */
        if (ialgo == ABO_ECMWF) free_ecmwf_state(s);
        else free_coare_state(s);
    }
    return ABO_OK;
}

/*
 * Station time series: the workflow of src/tests/test_aerobulk_buoy_series_oce.f90:364-537 (time loop) for S
 * independent stations.  Records are [Nt][S] (station index fastest).  Per record jt (1-based), per station:
 *   q_zt from dew-point / RH (:220-236; RH capped at 99.999 %), theta_zt = t_zt + gamma_moist*zt (:399),
 *   ssq = 0.98 q_sat(SST,SLP) (:413), Qsw = (1-albedo) rad_sw (:447), TURB_<algo>(kt=jt, l_use_cs = l_use_wl = l_skin)
 *   (:450-491; the program runs COARE 3.6 for the 'coare3p0' choice, SURVEY 8a quirk 8 -- a test-program slip that is
 *   NOT reproduced: 'coare3p0' runs TURB_COARE3P0), dT = Ts - SST (:493), t_zu by 4 lapse-rate passes (:499-503),
 *   RiB (:506), BULK_FORMULA (:509-512), Qlw, QNS (:515-518).
 * out[28] (NULL = skip), each [Nt][S]: rho_zu QL QH Qlw QNS Qsw dT_cs dT_wl TAU dT Hz_wl Qnt_ac Tau_ac Cd Ce Ch
 *   theta_zu q_zu t_zu RiB z0 u_star L UN10 Ts Evap q_zt theta_zt.
 * hum_kind: 0 specific humidity, 1 dew-point [K], 2 relative humidity [%].
 * The warm-layer state lives from record 1 to record Nt and is released on return (also after an error).
 */
int abo_series(abo_session *s, const char *calgo, int Nt, long S, double zt, double zu,
               const int *isecday_utc, const double *lon,
               const double *sst, const double *t_zt, const double *hum_zt, int hum_kind, const double *wnd,
               const double *slp, const double *rad_sw, const double *rad_lw, int l_skin, double *const *out)
{
    init_consts();
    int ialgo = algo_id(calgo);
    if (!ialgo) { snprintf(s->errmsg, sizeof(s->errmsg), "unknown algorithm %s", calgo); return ABO_ERR_ALGO; }
    int skin_algo = (ialgo == ABO_COARE3P0 || ialgo == ABO_COARE3P6 || ialgo == ABO_ECMWF);
    int lsk = (l_skin && skin_algo) ? 1 : 0;
    if (S <= 0 || Nt <= 0) return ABO_OK;
    const int nitend_saved = s->nitend;
    s->nitend = -1;   /* the state is released after the last record, below (its values are still reported) */
    double *w[17];
    for (int k = 0; k < 17; k++) w[k] = (double *)malloc((size_t)S * sizeof(double));
    double *Ts = w[0], *qs = w[1], *tha = w[2], *qa = w[3], *Qsw = w[4], *Cd = w[5], *Ch = w[6], *Ce = w[7],
           *thu = w[8], *qu = w[9], *Ub = w[10], *z0 = w[11], *us = w[12], *xL = w[13], *un10 = w[14],
           *dTcs = w[15], *dTwl = w[16];
    double *Hwl = (double *)malloc((size_t)S * sizeof(double));
    int rc = ABO_OK;
    for (int jt = 1; jt <= Nt && rc == ABO_OK; jt++) {
        const long o = (long)(jt - 1) * S;
        for (long i = 0; i < S; i++) {
            const double T = t_zt[o + i], P = slp[o + i];
            double q;
            if (hum_kind == 2) q = q_air_rh(MIN(99.999, hum_zt[o + i]), T, P);
            else if (hum_kind == 1) q = q_air_dp(hum_zt[o + i], P);
            else q = hum_zt[o + i];
            qa[i] = q;
            tha[i] = T + gamma_moist(T, q) * zt;
            qs[i] = rdct_qsat_salt * q_sat(sst[o + i], P);
            Ts[i] = sst[o + i];
            Qsw[i] = (1. - roce_alb0) * rad_sw[o + i];
            dTcs[i] = 0.; dTwl[i] = 0.; Hwl[i] = 0.;
        }
        double *opt[10] = {NULL, NULL, NULL, z0, us, xL, un10, dTcs, dTwl, Hwl};
        rc = abo_turb(s, calgo, jt, zt, zu, S, Ts, tha, qs, qa, wnd + o, lsk, lsk, Cd, Ch, Ce, thu, qu, Ub,
                      Qsw, rad_lw + o, slp + o, isecday_utc[jt - 1], lon, opt);
        if (rc != ABO_OK) break;
        for (long i = 0; i < S; i++) {
            double tz = thu[i];
            for (int jq = 0; jq < 4; jq++) tz = thu[i] - gamma_moist(tz, qu[i]) * zu;
            const double rib = Ri_bulk(zu, Ts[i], thu[i], qs[i], qu[i], Ub[i]);
            double tau, qh, ql, ev, rho;
            bulk_formula(zu, Ts[i], qs[i], thu[i], qu[i], Cd[i], Ch[i], Ce[i], wnd[o + i], Ub[i], slp[o + i],
                         &tau, &qh, &ql, &ev, &rho);
            if (tau > 10.) {   /* BULK_FORMULA_VCTR, mod_phymbl.f90:1250-1253 */
                snprintf(s->errmsg, sizeof(s->errmsg), "wind stress too strong (record %d, station %ld)", jt, i);
                rc = ABO_ERR_TAU;
            }
            const double qlw = qlw_net(rad_lw[o + i], Ts[i]);
            const double v[28] = {rho, ql, qh, qlw, qh + ql + qlw, Qsw[i], dTcs[i], dTwl[i], tau, Ts[i] - sst[o + i],
                                  Hwl[i],
                                  (lsk && ialgo != ABO_ECMWF && s->c_Qnt_ac) ? s->c_Qnt_ac[i] : 0.,
                                  (lsk && ialgo != ABO_ECMWF && s->c_Tau_ac) ? s->c_Tau_ac[i] : 0.,
                                  Cd[i], Ce[i], Ch[i], thu[i], qu[i], tz, rib, z0[i], us[i], xL[i], un10[i], Ts[i], ev,
                                  qa[i], tha[i]};
            for (int k = 0; k < 28; k++)
                if (out[k]) out[k][o + i] = v[k];
        }
    }
    if (lsk) {
        if (ialgo == ABO_ECMWF) free_ecmwf_state(s);
        else free_coare_state(s);
    }
    s->nitend = nitend_saved;
    for (int k = 0; k < 17; k++) free(w[k]);
    free(Hwl);
    return rc;
}

/* ================================================================== */
/* SEA ICE: src/ice/mod_blk_ice_{nemo,easy,an05,lu12,lg15,lg15_io}.f90,                   */
/* src/ice/mod_cdn_form_ice.f90 and the ice helpers of src/mod_phymbl.f90.                                        */
/* Not restated: mod_blk_ice_best.f90 (reads sqrtCdn10 before it is ever set, :154 -- undefined in the reference) */
/* and the over-water outputs of turb_ice_lg15_io (zCdN_s(:,:,2), zz0_s(:,:,2) are never initialised, :170-172).  */
/* ================================================================== */
static const double rtt0 = 273.16;            /* mod_const.f90:61 triple point */
static const double rLsub = 2.834e+6;         /* :92 */
static const double rCd_ice = 1.4e-3;         /* :118 */
static const double wspd_thrshld_ice = 0.2;   /* :120 */

/* e_sat_ice_sclr, mod_phymbl.f90:815-830 (Goff over ice) */
static double e_sat_ice(double pTa)
{
    const double rAg_i = -9.09718, rBg_i = -3.56654, rCg_i = 0.876793;
    const double rDg_i = 0.7858350313586662;    /* LOG10(6.1071_wp): a PARAMETER, folded (correctly rounded) by the compiler, :147 */
    double zta = MAX(pTa, 180.);
    double ztmp = rtt0 / zta;
    double zle = rAg_i * (ztmp - 1.) + rBg_i * log10(ztmp) + rCg_i * (1. - zta / rtt0) + rDg_i;
    return 100. * pow(10., zle);
}
/* q_sat_sclr with l_ice=.TRUE., :881-904 */
static double q_sat_ice(double pTa, double pslp)
{
    double ze_s = e_sat_ice(pTa);
    return reps0 * ze_s / (pslp - (1. - reps0) * ze_s);
}
/* Cd_from_z0 without ppsi, :1396-1414 */
static double Cd_from_z0(double pzu, double pz0)
{
    double r = 1. / log(pzu / pz0);
    return vkarmn2 * r * r;
}
/* f_m_louis_sclr :1419-1440, f_h_louis_sclr :1458-1479 (Louis 1979; rc_louis = 5, :149-153) */
static double f_m_louis(double pzu, double pRib, double pCdn, double pz0)
{
    const double rc2_louis = 5. * 5., ram_louis = 2. * 5.;
    double zstab = 0.5 + SIGN(0.5, pRib);
    double ztu = pRib / (1. + 3. * rc2_louis * pCdn * sqrt(fabs(-pRib * (pzu / pz0 + 1.))));
    double zts = pRib / sqrt(fabs(1. + pRib));
    return (1. - zstab) * (1. - ram_louis * ztu) + zstab * 1. / (1. + ram_louis * zts);
}
static double f_h_louis(double pzu, double pRib, double pChn, double pz0)
{
    const double rc2_louis = 5. * 5., rah_louis = 3. * 5.;
    double zstab = 0.5 + SIGN(0.5, pRib);
    double ztu = pRib / (1. + 3. * rc2_louis * pChn * sqrt(fabs(-pRib * (pzu / pz0 + 1.))));
    double zts = pRib / sqrt(fabs(1. + pRib));
    return (1. - zstab) * (1. - rah_louis * ztu) + zstab * 1. / (1. + rah_louis * zts);
}
/* CdN10_f_LU13, mod_cdn_form_ice.f90:170-208: rCe_0 A**(mu-1) (1-A)**(nu + 1/(10 beta)), mu = nu = 1, beta = 1.4 */
static double CdN10_f_LU13(double pfrice)
{
    const double rCe_0 = 2.23E-3, rNu_0 = 1., rMu_0 = 1., rbeta_0 = 1.4;
    double zcoef = rNu_0 + 1. / (10. * rbeta_0);
    return rCe_0 * pow(pfrice, rMu_0 - 1.) * pow(1. - pfrice, zcoef);
}
/* CdN_f_LG15_light, :299-330 (Eq.46).  NB: the reference assigns the WHOLE result array inside its point loop
 * (`CdN_f_LG15_light(:,:) = ...`, :324), so every point ends up with the value of the LAST point (Ni,Nj): callers
 * pass the ice fraction of the last point (see abo_turb_ice). */
static double CdN_f_LG15_light(double pzu, double pfrice, double pz0w)
{
    const double rce10_i_0 = 3.46e-3, rbeta_0 = 1.4;
    double ztmp = 1. / pz0w;
    double zrlog = log(10. * ztmp) / log(pzu * ztmp);
    return rce10_i_0 * zrlog * zrlog * pfrice * pow(1. - pfrice, rbeta_0);
}
/* psi_m_ice / psi_h_ice, mod_blk_ice_an05.f90:329-405 (same text in mod_blk_ice_easy.f90:213-289) */
static double psi_m_ice(double pzeta)
{
    double zta = pzeta;
    double zx = pow(fabs(1. - 16. * zta), .25);
    double zpsi_u = log((1. + zx * zx) / 2.) + 2. * log((1. + zx) / 2.) - 2. * atan(zx) + 0.5 * rpi;
    double zpsi_s = -(0.7 * zta + 0.75 * (zta - 14.3) * exp(-0.35 * zta) + 10.7);
    double zstab = 0.5 + SIGN(0.5, zta);
    return (1. - zstab) * zpsi_u + zstab * zpsi_s;
}
static double psi_h_ice(double pzeta)
{
    double zta = pzeta;
    double zx = pow(fabs(1. - 16. * zta), .25);
    double zpsi_u = 2. * log((1. + zx * zx) / 2.);
    double zpsi_s = -(0.7 * zta + 0.75 * (zta - 14.3) * exp(-0.35 * zta) + 10.7);
    double zstab = 0.5 + SIGN(0.5, zta);
    return (1. - zstab) * zpsi_u + zstab * zpsi_s;
}
/* rough_leng_m, mod_blk_ice_an05.f90:247-268 (Andreas et al. 2005 Eq.19) */
static double rough_leng_m(double pus, double pnua)
{
    double zus = MAX(pus, 1.E-9);
    double zz = (zus - 0.18) / 0.1;
    return 0.135 * pnua / zus + 0.035 * zus * zus / grav * (5. * exp(-zz * zz) + 1.);
}
/* rough_leng_tq, :270-325 (Andreas 1987 table); returns 1 when the reference would ctl_stop (:296-297: for
 * 2.49999 < R* < 2.5 none of the three regimes is selected) */
static int rough_leng_tq(double pz0, double pus, double pnua, double *z0t, double *z0q)
{
    double zz0 = pz0;
    double zus = MAX(pus, 1.E-9);
    double zre = MAX(zus * zz0 / pnua, 0.);
    double zsmoot = 0.5 + SIGN(0.5, (0.135 - zre));
    double ztrans = 0.5 + SIGN(0.5, (2.49999 - zre)) - zsmoot;
    double zrough = 0.5 + SIGN(0.5, (zre - 2.5));
    int bad = (zsmoot + ztrans + zrough > 1.001) || (zsmoot + ztrans + zrough < 0.999);
    double zlog = log(zre);
    double zlog2 = zlog * zlog;
    double zb0 = zsmoot * 1.25 + ztrans * 0.149 + zrough * 0.317;
    double zb1 = -ztrans * 0.550 - zrough * 0.565;
    double zb2 = -zrough * 0.183;
    *z0t = zz0 * exp(zb0 + zb1 * zlog + zb2 * zlog2);
    zb0 = zsmoot * 1.61 + ztrans * 0.351 + zrough * 0.396;
    zb1 = -ztrans * 0.628 - zrough * 0.512;
    zb2 = -zrough * 0.180;
    *z0q = zz0 * exp(zb0 + zb1 * zlog + zb2 * zlog2);
    return bad;
}

typedef struct {
    double Cd, Ch, Ce, t_zu, q_zu, Ub;
    double CdN, ChN, CeN, z0, us, L, UN10, CdN_frm;
} ice_out;

static void ice_first_guess(double Ts_i, double t_zt, double qs_i, double q_zt, double U_zu, ice_out *o, double *dt, double *dq)
{
    o->Ub = MAX(U_zu, wspd_thrshld_ice);
    o->t_zu = MAX(t_zt, 100.);
    o->q_zu = MAX(q_zt, 0.1e-6);
    *dt = o->t_zu - Ts_i; *dt = SIGN(MAX(fabs(*dt), 1.E-6), *dt);
    *dq = o->q_zu - qs_i; *dq = SIGN(MAX(fabs(*dq), 1.E-9), *dq);
}

/* turb_ice_nemo, mod_blk_ice_nemo.f90:36-153 */
static void turb_ice_nemo(double zt, double zu, double Ts_i, double t_zt, double qs_i, double q_zt, double U_zu, ice_out *o)
{
    (void)zt;
    double dt, dq;
    ice_first_guess(Ts_i, t_zt, qs_i, q_zt, U_zu, o, &dt, &dq);
    o->Cd = o->Ch = o->Ce = rCd_ice;
    o->CdN = o->ChN = o->CeN = rCd_ice;
    o->z0 = z0_from_Cd_neutral(zu, o->Cd);
    o->us = sqrt(rCd_ice) * o->Ub;
    o->L = 1. / One_on_L(o->t_zu, o->q_zu, sqrt(rCd_ice) * o->Ub, rCd_ice / sqrt(rCd_ice) * dt, rCd_ice / sqrt(rCd_ice) * dq);
    o->UN10 = sqrt(rCd_ice) * o->Ub / vkarmn * log(10. / z0_from_Cd_neutral(zu, o->Cd));
    o->CdN_frm = 0.;
}

/* turb_ice_easy, mod_blk_ice_easy.f90:35-209 (CdN, ChN, CeN are scalar INPUTS) */
static void turb_ice_easy(int nb_iter, double zt, double zu, double Ts_i, double t_zt, double qs_i, double q_zt, double U_zu,
                          double CdN, double ChN, double CeN, ice_out *o)
{
    int l_zt_equal_zu = (fabs(zu - zt) < 0.01);
    double zsqrtCDN = sqrt(CdN);
    double zlog1 = log(zt / zu);
    double zlog2 = log(zu / 10.);
    o->Ub = MAX(U_zu, wspd_thrshld_ice);
    o->t_zu = MAX(t_zt, 100.);
    o->q_zu = MAX(q_zt, 0.1e-6);
    o->Cd = CdN; o->Ch = ChN; o->Ce = CeN;
    double u_star = 0., t_star = 0., q_star = 0., zeta_u = 0., zeta_t = 0.;
    for (int jit = 1; jit <= nb_iter; jit++) {
        double dt_zu = o->t_zu - Ts_i;
        double dq_zu = o->q_zu - qs_i;
        double ztmp0 = sqrt(o->Cd);
        u_star = ztmp0 * o->Ub;
        ztmp0 = 1. / MAX(ztmp0, 1.E-15);
        t_star = o->Ch * dt_zu * ztmp0;
        q_star = o->Ce * dq_zu * ztmp0;
        ztmp0 = One_on_L(o->t_zu, o->q_zu, u_star, t_star, q_star);
        ztmp0 = SIGN(MIN(fabs(ztmp0), 200.), ztmp0);
        zeta_u = zu * ztmp0;
        zeta_u = SIGN(MIN(fabs(zeta_u), 50.0), zeta_u);
        if (!l_zt_equal_zu) {
            zeta_t = zt * ztmp0;
            zeta_t = SIGN(MIN(fabs(zeta_t), 50.0), zeta_t);
        }
        ztmp0 = 1. + zsqrtCDN / vkarmn * (zlog2 - psi_m_ice(zeta_u));
        o->Cd = MIN(MAX(CdN / (ztmp0 * ztmp0), Cx_min), 1.9E-3);
        ztmp0 = (zlog2 - psi_h_ice(zeta_u)) / vkarmn / zsqrtCDN;
        double ztmp1 = sqrt(o->Cd) / zsqrtCDN;
        o->Ch = MIN(MAX(ChN * ztmp1 / (1. + ChN * ztmp0), Cx_min), 1.9E-3);
        o->Ce = MIN(MAX(CeN * ztmp1 / (1. + CeN * ztmp0), Cx_min), 1.9E-3);
        if (!l_zt_equal_zu) {
            ztmp0 = psi_h_ice(zeta_u) - psi_h_ice(zeta_t) + zlog1;
            o->t_zu = t_zt - t_star / vkarmn * ztmp0;
            o->q_zu = MAX(0., q_zt - q_star / vkarmn * ztmp0);
        }
    }
    o->CdN = CdN; o->ChN = ChN; o->CeN = CeN;
    o->z0 = z0_from_Cd_psi(zu, o->Cd, psi_m_ice(zeta_u));
    o->us = u_star;
    o->L = 1. / One_on_L(o->t_zu, o->q_zu, u_star, t_star, q_star);
    o->UN10 = UN10_from_CD(zu, o->Ub, o->Cd, psi_m_ice(zeta_u));
    o->CdN_frm = 0.;
}

/* turb_ice_an05, mod_blk_ice_an05.f90:41-243; returns 1 when rough_leng_tq would ctl_stop */
static int turb_ice_an05(int nb_iter, double zt, double zu, double Ts_i, double t_zt, double qs_i, double q_zt, double U_zu, ice_out *o)
{
    int l_zt_equal_zu = (fabs(zu - zt) < 0.01);
    double dt_zu, dq_zu;
    int bad = 0;
    ice_first_guess(Ts_i, t_zt, qs_i, q_zt, U_zu, o, &dt_zu, &dq_zu);
    double znu_a = visc_air(o->t_zu);
    double z0 = 8.0E-4;
    double u_star = 0.035 * o->Ub * log(10. / z0) / log(zu / z0);
    z0 = rough_leng_m(u_star, znu_a);
    for (int jit = 1; jit <= 2; jit++) {
        u_star = MAX(o->Ub * vkarmn / (log(zu) - log(z0)), 1.E-9);
        z0 = rough_leng_m(u_star, znu_a);
    }
    double z0t, z0q;
    bad |= rough_leng_tq(z0, u_star, znu_a, &z0t, &z0q);
    double t_star = dt_zu * vkarmn / (log(zu / z0t));
    double q_star = dq_zu * vkarmn / (log(zu / z0q));
    for (int jit = 1; jit <= nb_iter; jit++) {
        double ztmp0 = One_on_L(o->t_zu, o->q_zu, u_star, t_star, q_star);
        ztmp0 = SIGN(MIN(fabs(ztmp0), 200.), ztmp0);
        double zeta_u = zu * ztmp0, zeta_t = 0.;
        zeta_u = SIGN(MIN(fabs(zeta_u), 50.0), zeta_u);
        if (!l_zt_equal_zu) {
            zeta_t = zt * ztmp0;
            zeta_t = SIGN(MIN(fabs(zeta_t), 50.0), zeta_t);
        }
        z0 = rough_leng_m(u_star, znu_a);
        bad |= rough_leng_tq(z0, u_star, znu_a, &z0t, &z0q);
        ztmp0 = psi_h_ice(zeta_u);
        t_star = dt_zu * vkarmn / (log(zu) - log(z0t) - ztmp0);
        q_star = dq_zu * vkarmn / (log(zu) - log(z0q) - ztmp0);
        u_star = MAX(o->Ub * vkarmn / (log(zu) - log(z0) - psi_m_ice(zeta_u)), 1.E-9);
        if (!l_zt_equal_zu) {
            double ztmp1 = log(zt / zu) + ztmp0 - psi_h_ice(zeta_t);
            o->t_zu = t_zt - t_star / vkarmn * ztmp1;
            o->q_zu = q_zt - q_star / vkarmn * ztmp1;
            dt_zu = o->t_zu - Ts_i; dt_zu = SIGN(MAX(fabs(dt_zu), 1.E-6), dt_zu);
            dq_zu = o->q_zu - qs_i; dq_zu = SIGN(MAX(fabs(dq_zu), 1.E-9), dq_zu);
        }
    }
    double ztmp0 = u_star / o->Ub;
    o->Cd = ztmp0 * ztmp0;
    o->Ch = ztmp0 * t_star / dt_zu;
    o->Ce = ztmp0 * q_star / dq_zu;
    ztmp0 = 1. / log(zu / z0);
    o->CdN = vkarmn2 * ztmp0 * ztmp0;
    o->ChN = vkarmn2 * ztmp0 / log(zu / z0t);
    o->CeN = vkarmn2 * ztmp0 / log(zu / z0q);
    o->z0 = z0;
    o->us = u_star;
    o->L = 1. / One_on_L(o->t_zu, o->q_zu, u_star, t_star, q_star);
    o->UN10 = u_star / vkarmn * log(10. / z0);
    o->CdN_frm = 0.;
    return bad;
}

/* turb_ice_lu12, mod_blk_ice_lu12.f90:50-214 ("Method #1": skin drag of z0 = 0.69e-3 m + LU13 form drag) */
static void turb_ice_lu12(double zt, double zu, double Ts_i, double t_zt, double qs_i, double q_zt, double U_zu, double frice, ice_out *o)
{
    (void)zt;
    const double rz0_i_s_0 = 0.69e-3;
    double dt_zu, dq_zu;
    ice_first_guess(Ts_i, t_zt, qs_i, q_zt, U_zu, o, &dt_zu, &dq_zu);
    o->CdN_frm = CdN10_f_LU13(frice);
    o->Cd = Cd_from_z0(zu, rz0_i_s_0) + o->CdN_frm;
    o->Ch = o->Cd;
    o->Ce = o->Cd;
    o->CdN = o->Cd; o->ChN = o->Ch; o->CeN = o->Ce;
    o->z0 = z0_from_Cd_neutral(zu, o->Cd);
    o->us = sqrt(o->Cd) * o->Ub;
    o->L = 1. / One_on_L(o->t_zu, o->q_zu, sqrt(o->Cd) * o->Ub, o->Cd / sqrt(o->Cd) * dt_zu, o->Cd / sqrt(o->Cd) * dq_zu);
    o->UN10 = sqrt(o->Cd) * o->Ub / vkarmn * log(10. / z0_from_Cd_neutral(zu, o->Cd));
}

/* turb_ice_lg15, mod_blk_ice_lg15.f90:53-307 (== the over-ice part of turb_ice_lg15_io, mod_blk_ice_lg15_io.f90:39-370).
 * frice_form is the ice fraction the form drag is computed from: the LAST point's in the reference (see CdN_f_LG15_light). */
static void turb_ice_lg15(int nb_iter, double zt, double zu, double Ts_i, double t_zt, double qs_i, double q_zt, double U_zu,
                          double frice_form, ice_out *o)
{
    const double ralpha_0 = 0.2, rz0_i_s_0 = 0.69e-3, rz0_i_f_0 = 4.54e-4;
    int l_zt_equal_zu = (fabs(zu - zt) < 0.01);
    double dt_zu, dq_zu;
    ice_first_guess(Ts_i, t_zt, qs_i, q_zt, U_zu, o, &dt_zu, &dq_zu);
    double zz0_s = rz0_i_s_0;
    double zCdN_s = Cd_from_z0(zu, zz0_s);
    double zChN_s = vkarmn2 / (log(zu / zz0_s) * log(zu / (ralpha_0 * zz0_s)));
    double zz0_f = rz0_i_f_0;
    double zCdN_f = CdN_f_LG15_light(zu, frice_form, zz0_f);
    double zChN_f = zCdN_f / (1. + log(1. / ralpha_0) / vkarmn * sqrt(zCdN_f));
    o->Cd = zCdN_s + zCdN_f;
    o->Ch = zChN_s + zChN_f;
    double RiB = Ri_bulk(zt, Ts_i, t_zt, qs_i, q_zt, o->Ub);
    for (int jit = 1; jit <= nb_iter; jit++) {
        double xtmp1, xtmp2;
        if (!l_zt_equal_zu) {
            xtmp1 = zCdN_s + zCdN_f;
            xtmp2 = zz0_s + zz0_f;
            xtmp1 = log(zt / zu) + f_h_louis(zu, RiB, xtmp1, xtmp2) - f_h_louis(zt, RiB, xtmp1, xtmp2);
            xtmp2 = MAX(o->Ub + (sqrt(o->Cd) * o->Ub) * xtmp1, wspd_thrshld_ice);
            xtmp2 = MIN(xtmp2, o->Ub);
        } else {
            xtmp2 = o->Ub;
        }
        RiB = Ri_bulk(zt, Ts_i, t_zt, qs_i, q_zt, xtmp2);
        o->Cd = zCdN_s * f_m_louis(zu, RiB, zCdN_s, zz0_s);
        o->Ch = zChN_s * f_h_louis(zu, RiB, zCdN_s, zz0_s);
        o->Cd = o->Cd + zCdN_f * f_m_louis(zu, RiB, zCdN_f, zz0_f);
        o->Ch = o->Ch + zChN_f * f_h_louis(zu, RiB, zCdN_f, zz0_f);
        if (!l_zt_equal_zu) {
            xtmp1 = zCdN_s + zCdN_f;
            xtmp2 = zz0_s + zz0_f;
            xtmp1 = log(zt / zu) + f_h_louis(zu, RiB, xtmp1, xtmp2) - f_h_louis(zt, RiB, xtmp1, xtmp2);
            xtmp2 = 1. / sqrt(o->Cd);
            o->t_zu = t_zt - (o->Ch * dt_zu * xtmp2) / vkarmn * xtmp1;
            o->q_zu = q_zt - (o->Ch * dq_zu * xtmp2) / vkarmn * xtmp1;
            o->q_zu = MAX(0., o->q_zu);
            dt_zu = o->t_zu - Ts_i;
            dq_zu = o->q_zu - qs_i;
            dt_zu = SIGN(MAX(fabs(dt_zu), 1.E-6), dt_zu);
            dq_zu = SIGN(MAX(fabs(dq_zu), 1.E-9), dq_zu);
        }
    }
    o->Ce = o->Ch;
    o->CdN_frm = zCdN_f;
    o->CdN = zCdN_s + zCdN_f;
    o->ChN = zChN_s + zChN_f;
    o->CeN = zChN_s + zChN_f;
    o->z0 = z0_from_Cd_neutral(zu, zCdN_s + zCdN_f);
    o->us = sqrt(o->Cd) * o->Ub;
    {
        double x = sqrt(o->Cd);
        o->L = 1. / One_on_L(o->t_zu, o->q_zu, x * o->Ub, o->Ch * dt_zu / x, o->Ce * dq_zu / x);
    }
    o->UN10 = sqrt(o->Cd) * o->Ub / vkarmn * log(10. / z0_from_Cd_neutral(zu, zCdN_s + zCdN_f));
}

enum { ABO_ICE_NEMO = 1, ABO_ICE_EASY = 2, ABO_ICE_AN05 = 3, ABO_ICE_LU12 = 4, ABO_ICE_LG15 = 5, ABO_ICE_LG15_IO = 6 };
static int ice_algo_id(const char *c)
{
    if (!strcmp(c, "nemo")) return ABO_ICE_NEMO;
    if (!strcmp(c, "easy")) return ABO_ICE_EASY;
    if (!strcmp(c, "an05")) return ABO_ICE_AN05;
    if (!strcmp(c, "lu12")) return ABO_ICE_LU12;
    if (!strcmp(c, "lg15")) return ABO_ICE_LG15;
    if (!strcmp(c, "lg15_io")) return ABO_ICE_LG15_IO;
    return 0;
}

static int ice_point(int ialgo, int nb_iter, double zt, double zu, double Ts_i, double t_zt, double qs_i, double q_zt,
                     double U_zu, double frice, double frice_form, const double *cxn, ice_out *o)
{
    switch (ialgo) {
    case ABO_ICE_NEMO: turb_ice_nemo(zt, zu, Ts_i, t_zt, qs_i, q_zt, U_zu, o); return 0;
    case ABO_ICE_EASY: turb_ice_easy(nb_iter, zt, zu, Ts_i, t_zt, qs_i, q_zt, U_zu, cxn[0], cxn[1], cxn[2], o); return 0;
    case ABO_ICE_AN05: return turb_ice_an05(nb_iter, zt, zu, Ts_i, t_zt, qs_i, q_zt, U_zu, o);
    case ABO_ICE_LU12: turb_ice_lu12(zt, zu, Ts_i, t_zt, qs_i, q_zt, U_zu, frice, o); return 0;
    default: turb_ice_lg15(nb_iter, zt, zu, Ts_i, t_zt, qs_i, q_zt, U_zu, frice_form, o); return 0;
    }
}

/*
 * Direct TURB_ICE_* call on n points.  calgo: nemo | easy | an05 | lu12 | lg15 | lg15_io.  frice is needed by lu12 / lg15 /
 * lg15_io; cxn[3] = CdN, ChN, CeN are the scalar inputs of `easy`.  per_point_form_drag = 0 reproduces the reference
 * (lg15: the form drag of EVERY point comes from the ice fraction of the LAST point, see CdN_f_LG15_light), 1 uses each
 * point's own fraction.  opt[8] (NULL = skip): CdN ChN CeN xz0 xu_star xL xUN10 CdN_frm.
 */
int abo_turb_ice(abo_session *s, const char *calgo, double zt, double zu, long n,
                 const double *Ts_i, const double *t_zt, const double *qs_i, const double *q_zt, const double *U_zu,
                 const double *frice, const double *cxn, int per_point_form_drag,
                 double *Cd, double *Ch, double *Ce, double *t_zu, double *q_zu, double *Ubzu, double *const *opt)
{
    init_consts();
    s->errmsg[0] = 0;
    int ialgo = ice_algo_id(calgo);
    if (!ialgo) { snprintf(s->errmsg, sizeof(s->errmsg), "unknown sea-ice algorithm %s", calgo); return ABO_ERR_ALGO; }
    if (ialgo >= ABO_ICE_LU12 && !frice) { snprintf(s->errmsg, sizeof(s->errmsg), "turb_ice_%s needs frice", calgo); return ABO_ERR_ALGO; }
    if (ialgo == ABO_ICE_EASY && !cxn) { snprintf(s->errmsg, sizeof(s->errmsg), "turb_ice_easy needs CdN, ChN, CeN"); return ABO_ERR_ALGO; }
    long bad = n;   /* first point where rough_leng_tq would ctl_stop */
    const int nb_iter = s->nb_iter;
#ifdef _OPENMP
#pragma omp parallel for num_threads(s->nthreads) schedule(static) reduction(min:bad)
#endif
    for (long i = 0; i < n; i++) {
        ice_out o;
        memset(&o, 0, sizeof(o));
        double fr = frice ? frice[i] : 0.;
        double frf = frice ? (per_point_form_drag ? frice[i] : frice[n - 1]) : 0.;
        if (ice_point(ialgo, nb_iter, zt, zu, Ts_i[i], t_zt[i], qs_i[i], q_zt[i], U_zu[i], fr, frf, cxn, &o) && i < bad) bad = i;
        Cd[i] = o.Cd; Ch[i] = o.Ch; Ce[i] = o.Ce; t_zu[i] = o.t_zu; q_zu[i] = o.q_zu; Ubzu[i] = o.Ub;
        if (opt) {
            const double v[8] = {o.CdN, o.ChN, o.CeN, o.z0, o.us, o.L, o.UN10, o.CdN_frm};
            for (int k = 0; k < 8; k++)
                if (opt[k]) opt[k][i] = v[k];
        }
    }
    if (bad < n) {
        snprintf(s->errmsg, sizeof(s->errmsg), " rough_leng_tq@mod_blk_ice_an05.f90 => something wrong with zsmoot, ztrans, zrough! (point %ld)", bad + 1);
        return ABO_ERR_ICE_ROUGH;
    }
    return ABO_OK;
}

/* BULK_FORMULA_SCLR with l_ice = .TRUE., mod_phymbl.f90:1149-1203 */
static void bulk_formula_ice(double pzu, double pts, double pqs, double pThta, double pqa, double pCd, double pCh, double pCe,
                             double pwnd, double pUb, double pslp, double *pTau, double *pQsen, double *pQlat, double *pEvap)
{
    double zta = pThta - rgamma_dry * pzu;
    double zrho = rho_air(zta, pqa, pslp);
    zrho = rho_air(zta, pqa, pslp - zrho * grav * pzu);
    double zUrho = pUb * MAX(zrho, 1.);
    *pTau = zUrho * pCd * pwnd;
    double zevap = zUrho * pCe * (pqa - pqs);
    *pQsen = zUrho * pCh * (pThta - pts) * cp_air(pqa);
    *pQlat = rLsub * zevap;
    *pEvap = MIN(zevap, 0.);
}

/*
 * Ice + leads workflow of src/ice/test_aerobulk_oce+ice.f90:225-412 (and test_aerobulk_ice.f90:186-370) on n points:
 *   siq = q_sat(SIT, SLP, l_ice) (:212); ssq = 0.98 q_sat(SST, SLP) (the program evaluates it at SIT, :213 -- a slip of
 *   the test program that is not reproduced); theta_zt = t_zt + gamma_moist zt (:262); over the leads
 *   TURB_<calgo_oce>(no skin) + BULK_FORMULA (:297-304); over the ice TURB_ICE_<calgo_ice> (:332-351), Ri_b (:354),
 *   t_zu by 4 lapse-rate passes at the mean layer temperature (:359-363), rho_zu (:384-387), BULK_FORMULA(l_ice) (:391).
 * hum_kind 0 q, 1 dew-point [K], 2 RH [%].  out[35] (NULL = skip):
 *   ice  0 Cd 1 Ch 2 Ce 3 theta_zu 4 q_zu 5 t_zu 6 Ub 7 RiB 8 z0 9 u* 10 L 11 UN10 12 rho_zu 13 Tau 14 QH 15 QL 16 Evap
 *   water 17 Cd 18 Ch 19 Ce 20 theta_zu 21 q_zu 22 Ub 23 z0 24 u* 25 L 26 UN10 27 Tau 28 QH 29 QL 30 Evap
 *   cell  31 Tau 32 QH 33 QL 34 Evap  = A ice + (1-A) water  (NEMO-style partition; not in the reference program)
 */
int abo_oce_ice(abo_session *s, const char *calgo_ice, const char *calgo_oce, double zt, double zu, long n,
                const double *sit, const double *sst, const double *t_zt, const double *hum_zt, int hum_kind,
                const double *wnd, const double *slp, const double *frice, const double *cxn, int per_point_form_drag,
                double *const *out)
{
    init_consts();
    s->errmsg[0] = 0;
    int ialgo = ice_algo_id(calgo_ice);
    if (!ialgo) { snprintf(s->errmsg, sizeof(s->errmsg), "unknown sea-ice algorithm %s", calgo_ice); return ABO_ERR_ALGO; }
    int ioce = calgo_oce ? algo_id(calgo_oce) : 0;
    if (calgo_oce && !ioce) { snprintf(s->errmsg, sizeof(s->errmsg), "unknown algorithm %s", calgo_oce); return ABO_ERR_ALGO; }
    if (ialgo == ABO_ICE_EASY && !cxn) { snprintf(s->errmsg, sizeof(s->errmsg), "turb_ice_easy needs CdN, ChN, CeN"); return ABO_ERR_ALGO; }
    long bad = n;
    int badtau = 0;
    const int nb_iter = s->nb_iter;
#ifdef _OPENMP
#pragma omp parallel for num_threads(s->nthreads) schedule(static) reduction(min:bad) reduction(|:badtau)
#endif
    for (long i = 0; i < n; i++) {
        const double T = t_zt[i], P = slp[i];
        double q;
        if (hum_kind == 2) q = q_air_rh(hum_zt[i], T, P);
        else if (hum_kind == 1) q = q_air_dp(hum_zt[i], P);
        else q = hum_zt[i];
        const double siq = q_sat_ice(sit[i], P);
        const double tha = T + gamma_moist(T, q) * zt;
        double v[35];
        for (int k = 0; k < 35; k++) v[k] = 0.;
        ice_out o;
        memset(&o, 0, sizeof(o));
        if (ice_point(ialgo, nb_iter, zt, zu, sit[i], tha, siq, q, wnd[i], frice[i],
                      per_point_form_drag ? frice[i] : frice[n - 1], cxn, &o) && i < bad) bad = i;
        double tz = o.t_zu;
        for (int jq = 0; jq < 4; jq++) tz = o.t_zu - gamma_moist(0.5 * (tz + sit[i]), o.q_zu) * zu;
        double rho = rho_air(tz, o.q_zu, P);
        rho = rho_air(tz, o.q_zu, P - rho * grav * zu);
        double tau, qh, ql, ev;
        bulk_formula_ice(zu, sit[i], siq, o.t_zu, o.q_zu, o.Cd, o.Ch, o.Ce, wnd[i], o.Ub, P, &tau, &qh, &ql, &ev);
        if (tau > 10.) badtau = 1;
        v[0] = o.Cd; v[1] = o.Ch; v[2] = o.Ce; v[3] = o.t_zu; v[4] = o.q_zu; v[5] = tz; v[6] = o.Ub;
        v[7] = Ri_bulk(zu, sit[i], o.t_zu, siq, o.q_zu, o.Ub);
        v[8] = o.z0; v[9] = o.us; v[10] = o.L; v[11] = o.UN10; v[12] = rho; v[13] = tau; v[14] = qh; v[15] = ql; v[16] = ev;
        if (ioce) {
            turb_io w;
            memset(&w, 0, sizeof(w));
            w.T_s = sst[i];
            w.q_s = rdct_qsat_salt * q_sat(sst[i], P);
            skin_in sk;
            memset(&sk, 0, sizeof(sk));
            switch (ioce) {
            case ABO_COARE3P0: turb_coare3p0(nb_iter, zt, zu, tha, q, wnd[i], &sk, &w); break;
            case ABO_COARE3P6: turb_coare3p6(nb_iter, zt, zu, tha, q, wnd[i], &sk, &w); break;
            case ABO_NCAR: turb_ncar(nb_iter, zt, zu, w.T_s, tha, w.q_s, q, wnd[i], &w); break;
            case ABO_ECMWF: turb_ecmwf(nb_iter, zt, zu, tha, q, wnd[i], &sk, &w); break;
            default: turb_andreas(nb_iter, zt, zu, w.T_s, tha, w.q_s, q, wnd[i], &w); break;
            }
            double tw, qhw, qlw, evw;
            bulk_formula(zu, sst[i], w.q_s, w.t_zu, w.q_zu, w.Cd, w.Ch, w.Ce, wnd[i], w.Ubzu, P, &tw, &qhw, &qlw, &evw, NULL);
            if (tw > 10.) badtau = 1;
            v[17] = w.Cd; v[18] = w.Ch; v[19] = w.Ce; v[20] = w.t_zu; v[21] = w.q_zu; v[22] = w.Ubzu; v[23] = w.z0;
            v[24] = w.us; v[25] = w.L; v[26] = w.UN10; v[27] = tw; v[28] = qhw; v[29] = qlw; v[30] = evw;
            const double A = frice[i];
            v[31] = A * tau + (1. - A) * tw; v[32] = A * qh + (1. - A) * qhw;
            v[33] = A * ql + (1. - A) * qlw; v[34] = A * ev + (1. - A) * evw;
        }
        for (int k = 0; k < 35; k++)
            if (out[k]) out[k][i] = v[k];
    }
    if (bad < n) {
        snprintf(s->errmsg, sizeof(s->errmsg), " rough_leng_tq@mod_blk_ice_an05.f90 => something wrong with zsmoot, ztrans, zrough! (point %ld)", bad + 1);
        return ABO_ERR_ICE_ROUGH;
    }
    if (badtau) { snprintf(s->errmsg, sizeof(s->errmsg), "wind stress too strong"); return ABO_ERR_TAU; }
    return ABO_OK;
}

/*
 * Sea-ice station series: the per-record computation of src/ice/test_aerobulk_buoy_series_ice.f90:326-470 on n records
 * (no state is carried between records; the program's nx = ny = 1 makes every TURB_ICE_* call a one-point call, so the
 * LG15 form drag uses the record's own ice concentration).
 *   rho_zt .. (:350, not returned); SIQ = q_sat(SIT, SLP, l_ice) (:362); theta_zt = t_zt + gamma_moist(t_zt, q_zt) zt (:371);
 *   RiB_zt = Ri_bulk(zt, SIT, theta_zt, SIQ, q_zt, MAX(W10, wspd_thrshld_ice)) (:380); Qsw = (1 - rice_alb0) rad_sw (:384);
 *   only where SIC > 0.01 (:388): TURB_ICE_<nemo|an05|lu12|lg15> (:392-411), RiB_zu (:427),
 *   BULK_FORMULA(l_ice, pEvap, prhoa) (:430-433), Qlw = qlw_net(rad_lw, SIT, l_ice) (:436), QNS = QH + QL + Qlw (:439).
 * Records without ice are skipped by the program (its arrays keep whatever ALLOCATE left there); here they read 0.
 * hum_kind 0 q, 1 dew-point [K], 2 RH [%] (:194-207).  out[21] (NULL = skip):
 *   0 rho_zu 1 QL 2 QH 3 Qlw 4 QNS 5 Qsw 6 TAU 7 SBLM [kg/m^2/s] 8 Cd_i 9 Ch_i 10 Ce_i 11 z0 12 RiB_zt 13 RiB_zu 14 CdN
 *   15 u_star 16 L 17 UN10 18 theta_zu 19 q_zu 20 Ublk
 */
static const double rice_alb0 = 0.8;    /* mod_const.f90:51 */
static const double emiss_i = 0.996;    /* mod_const.f90:56 */

int abo_series_ice(abo_session *s, const char *calgo, double zt, double zu, long n, const double *sic, const double *sit,
                   const double *t_zt, const double *hum_zt, int hum_kind, const double *wnd, const double *slp,
                   const double *rad_sw, const double *rad_lw, double *const *out)
{
    init_consts();
    s->errmsg[0] = 0;
    int ialgo = ice_algo_id(calgo);
    if (ialgo != ABO_ICE_NEMO && ialgo != ABO_ICE_AN05 && ialgo != ABO_ICE_LU12 && ialgo != ABO_ICE_LG15) {
        snprintf(s->errmsg, sizeof(s->errmsg), "UNKNOWN algo: %s !!!", calgo);   /* :413-415 */
        return ABO_ERR_ALGO;
    }
    long bad = n;
    int badtau = 0;
    const int nb_iter = s->nb_iter;
#ifdef _OPENMP
#pragma omp parallel for num_threads(s->nthreads) schedule(static) reduction(min:bad) reduction(|:badtau)
#endif
    for (long i = 0; i < n; i++) {
        const double T = t_zt[i], P = slp[i];
        double q;
        if (hum_kind == 2) q = q_air_rh(hum_zt[i], T, P);
        else if (hum_kind == 1) q = q_air_dp(hum_zt[i], P);
        else q = hum_zt[i];
        const double siq = q_sat_ice(sit[i], P);
        const double tha = T + gamma_moist(T, q) * zt;
        double v[21];
        for (int k = 0; k < 21; k++) v[k] = 0.;
        v[12] = Ri_bulk(zt, sit[i], tha, siq, q, MAX(wnd[i], wspd_thrshld_ice));
        v[5] = (1. - rice_alb0) * rad_sw[i];
        if (sic[i] > 0.01) {
            ice_out o;
            memset(&o, 0, sizeof(o));
            if (ice_point(ialgo, nb_iter, zt, zu, sit[i], tha, siq, q, wnd[i], sic[i], sic[i], NULL, &o) && i < bad) bad = i;
            double tau, qh, ql, ev;
            bulk_formula_ice(zu, sit[i], siq, o.t_zu, o.q_zu, o.Cd, o.Ch, o.Ce, wnd[i], o.Ub, P, &tau, &qh, &ql, &ev);
            if (tau > 10.) badtau = 1;
            /* prhoa of BULK_FORMULA: the density before its MAX(., 1) */
            const double zta = o.t_zu - rgamma_dry * zu;
            double rho = rho_air(zta, o.q_zu, P);
            rho = rho_air(zta, o.q_zu, P - rho * grav * zu);
            const double zt2 = sit[i] * sit[i];
            const double qlw = emiss_i * (rad_lw[i] - stefan * zt2 * zt2);   /* qlw_net_sclr with l_ice, mod_phymbl.f90:1306-1312 */
            v[0] = rho; v[1] = ql; v[2] = qh; v[3] = qlw; v[4] = qh + ql + qlw; v[6] = tau; v[7] = ev;
            v[8] = o.Cd; v[9] = o.Ch; v[10] = o.Ce; v[11] = o.z0;
            v[13] = Ri_bulk(zu, sit[i], o.t_zu, siq, o.q_zu, o.Ub);
            v[14] = o.CdN; v[15] = o.us; v[16] = o.L; v[17] = o.UN10; v[18] = o.t_zu; v[19] = o.q_zu; v[20] = o.Ub;
        }
        for (int k = 0; k < 21; k++)
            if (out[k]) out[k][i] = v[k];
    }
    if (bad < n) {
        snprintf(s->errmsg, sizeof(s->errmsg), " rough_leng_tq@mod_blk_ice_an05.f90 => something wrong with zsmoot, ztrans, zrough! (point %ld)", bad + 1);
        return ABO_ERR_ICE_ROUGH;
    }
    if (badtau) { snprintf(s->errmsg, sizeof(s->errmsg), "wind stress too strong"); return ABO_ERR_TAU; }
    return ABO_OK;
}

/* ------------------------------------------------------------------ */
/* building blocks for unit tests                                      */
/* ------------------------------------------------------------------ */
double abo_e_sat(double T) { init_consts(); return e_sat(T); }
double abo_q_sat(double T, double p) { init_consts(); return q_sat(T, p); }
double abo_theta_from_z_P0_T_q(double z, double slp, double T, double q) { init_consts(); return Theta_from_z_P0_T_q(z, slp, T, q); }
double abo_rho_air(double T, double q, double p) { init_consts(); return rho_air(T, q, p); }
double abo_visc_air(double T) { init_consts(); return visc_air(T); }
double abo_L_vap(double T) { init_consts(); return L_vap(T); }
double abo_cp_air(double q) { init_consts(); return cp_air(q); }
double abo_gamma_moist(double T, double q) { init_consts(); return gamma_moist(T, q); }
double abo_e_sat_ice(double T) { init_consts(); return e_sat_ice(T); }
double abo_q_sat_ice(double T, double p) { init_consts(); return q_sat_ice(T, p); }
double abo_f_m_louis(double zu, double Rib, double Cdn, double z0) { init_consts(); return f_m_louis(zu, Rib, Cdn, z0); }
double abo_f_h_louis(double zu, double Rib, double Chn, double z0) { init_consts(); return f_h_louis(zu, Rib, Chn, z0); }
double abo_psi_m_ice(double zeta) { init_consts(); return psi_m_ice(zeta); }
double abo_psi_h_ice(double zeta) { init_consts(); return psi_h_ice(zeta); }
double abo_rough_leng_m(double us, double nua) { init_consts(); return rough_leng_m(us, nua); }
int abo_rough_leng_tq(double z0, double us, double nua, double *z0t, double *z0q) { init_consts(); return rough_leng_tq(z0, us, nua, z0t, z0q); }
double abo_CdN10_f_LU13(double A) { init_consts(); return CdN10_f_LU13(A); }
double abo_CdN_f_LG15_light(double zu, double A, double z0w) { init_consts(); return CdN_f_LG15_light(zu, A, z0w); }
double abo_alpha_sw(double T) { init_consts(); return alpha_sw(T); }
double abo_qlw_net(double rlw, double Ts) { init_consts(); return qlw_net(rlw, Ts); }
double abo_one_on_L(double tha, double qa, double us, double ts, double qs) { init_consts(); return One_on_L(tha, qa, us, ts, qs); }
double abo_Ri_bulk(double z, double sst, double tha, double ssq, double qa, double ub) { init_consts(); return Ri_bulk(z, sst, tha, ssq, qa, ub); }
double abo_q_air_rh(double rh, double T, double p) { init_consts(); return q_air_rh(rh, T, p); }
double abo_q_air_dp(double dp, double p) { init_consts(); return q_air_dp(dp, p); }
double abo_z0tq_LKB(int iflag, double Rer, double z0) { init_consts(); return z0tq_LKB(iflag, Rer, z0); }
double abo_delta_skin_layer(double alpha, double Qd, double us, int has_qlat, double Qlat) { init_consts(); return delta_skin_layer(alpha, Qd, us, has_qlat, Qlat); }
double abo_cd_n10_ncar(double w) { init_consts(); return cd_n10_ncar(w); }
double abo_charn_coare3p0(double w) { init_consts(); return charn_coare3p0(w); }
double abo_charn_coare3p6(double w) { init_consts(); return charn_coare3p6(w); }
double abo_u_star_andreas(double u) { init_consts(); return u_star_andreas(u); }

double abo_psi_m(int algo, double zeta)
{
    init_consts();
    switch (algo) {
    case ABO_NCAR: return psi_m_ncar(zeta);
    case ABO_ECMWF: return psi_m_ecmwf(zeta);
    case ABO_ANDREAS: return psi_m_andreas(zeta);
    default: return psi_m_coare(zeta);
    }
}
double abo_psi_h(int algo, double zeta)
{
    init_consts();
    switch (algo) {
    case ABO_NCAR: return psi_h_ncar(zeta);
    case ABO_ECMWF: return psi_h_ecmwf(zeta);
    case ABO_ANDREAS: return psi_h_andreas(zeta);
    default: return psi_h_coare(zeta);
    }
}

void abo_turb_noskin(int algo, int nb_iter, double zt, double zu, double sst, double tha_zt, double ssq,
                     double q_zt, double U_zu, double *out13)
{
    init_consts();
    turb_io o;
    memset(&o, 0, sizeof(o));
    o.T_s = sst;
    o.q_s = ssq;
    skin_in sk;
    memset(&sk, 0, sizeof(sk));
    switch (algo) {
    case ABO_COARE3P0: turb_coare3p0(nb_iter, zt, zu, tha_zt, q_zt, U_zu, &sk, &o); break;
    case ABO_COARE3P6: turb_coare3p6(nb_iter, zt, zu, tha_zt, q_zt, U_zu, &sk, &o); break;
    case ABO_NCAR: turb_ncar(nb_iter, zt, zu, sst, tha_zt, ssq, q_zt, U_zu, &o); break;
    case ABO_ECMWF: turb_ecmwf(nb_iter, zt, zu, tha_zt, q_zt, U_zu, &sk, &o); break;
    default: turb_andreas(nb_iter, zt, zu, sst, tha_zt, ssq, q_zt, U_zu, &o); break;
    }
    out13[0] = o.Cd; out13[1] = o.Ch; out13[2] = o.Ce; out13[3] = o.t_zu; out13[4] = o.q_zu; out13[5] = o.Ubzu;
    out13[6] = o.CdN; out13[7] = o.ChN; out13[8] = o.CeN; out13[9] = o.z0; out13[10] = o.us; out13[11] = o.L;
    out13[12] = o.UN10;
}
