/*
 * aerobulk_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C restatement of the reference's `aerobulk_model` hot path
 * (brodeau/aerobulk, Fortran 90).  It exists only to CHECK the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  Nothing under aerobulk_b200/ links,
 * imports or calls it.
 *
 * Parity status: PINNED to the reference's own captured output
 * /root/reference/doc/ex_ab.dat (7 significant digits, 5 algorithms x 2
 * points, nb_iter=50) -- see tests/test_oracle_golden.py and
 * tests/golden/ex_ab.json.  The reference itself cannot be compiled here (no
 * Fortran compiler in the image), so warm-layer time integration (Rsw>0,
 * multi-step), rh/dp humidity inputs and zt==zu are defined by this
 * restatement alone ("parity unpinned" for those sub-paths).
 *
 * Arithmetic contract mirrored: gfortran -O2 -fdefault-real-8 on x86-64
 * (every REAL literal is FP64, no FMA contraction, libm pow/log/exp/atan,
 * SIGN == copysign, MODULO == floored modulo, INT == truncation).
 * Build with: gcc -O2 -ffp-contract=off (see oracle/Makefile).
 */
#ifndef AEROBULK_ORACLE_H
#define AEROBULK_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* algorithm ids (strings as in mod_aerobulk_compute.f90:129-176) */
enum { ABO_COARE3P0 = 1, ABO_COARE3P6 = 2, ABO_NCAR = 3, ABO_ECMWF = 4, ABO_ANDREAS = 5 };

/* error codes returned by abo_model (the reference STOPs; the oracle reports) */
enum {
    ABO_OK = 0,
    ABO_ERR_JT = 1,            /* mod_aerobulk.f90:244 */
    ABO_ERR_SKIN_ALGO = 2,     /* mod_aerobulk.f90:69-70 */
    ABO_ERR_SKIN_NORAD = 3,    /* mod_aerobulk.f90:72 */
    ABO_ERR_ALL_MASKED = 4,    /* mod_aerobulk.f90:122 */
    ABO_ERR_HUMIDITY = 5,      /* mod_phymbl.f90:1996-2003 */
    ABO_ERR_UNITS = 6,         /* mod_phymbl.f90:1946-1950 */
    ABO_ERR_ALGO = 7,          /* mod_aerobulk_compute.f90:173-176 */
    ABO_ERR_TAU = 8,           /* mod_phymbl.f90:1250-1253 */
    ABO_ERR_STATE = 9,         /* double ALLOCATE of warm-layer state, mod_blk_coare3p6.f90:82-83 */
    ABO_ERR_ICE_ROUGH = 10     /* rough_leng_tq ctl_stop, src/ice/mod_blk_ice_an05.f90:296-297 */
};

typedef struct abo_session abo_session;

abo_session *abo_new(void);
void abo_free(abo_session *s);

/* module globals of mod_const.f90:22-33 that callers may overwrite */
void abo_set_rdt(abo_session *s, double rdt);
void abo_set_gdept(abo_session *s, double gdept);
void abo_set_nb_iter(abo_session *s, int nb_iter);
int abo_get_nb_iter(const abo_session *s);
int abo_get_use_skin(const abo_session *s);
const char *abo_get_humidity_type(const abo_session *s);
void abo_set_threads(abo_session *s, int nthreads); /* row-block threads (the reference is serial) */
const char *abo_errmsg(const abo_session *s);

/*
 * AEROBULK_MODEL (mod_aerobulk.f90:176-269).  Arrays are (Ni,Nj) column-major,
 * contiguous.  Optional arguments: Niter / l_use_skin are NULL when absent;
 * rad_sw, rad_lw, T_s are NULL when absent.
 */
int abo_model(abo_session *s, int jt, int Nt, const char *calgo, double zt, double zu,
              int Ni, int Nj,
              const double *sst, const double *t_zt, const double *hum_zt,
              const double *U_zu, const double *V_zu, const double *slp,
              double *QL, double *QH, double *Tau_x, double *Tau_y, double *Evap,
              const int *Niter, const int *l_use_skin,
              const double *rad_sw, const double *rad_lw, double *T_s);

/* copy of the persistent warm-layer state (which: 0 dT_wl, 1 Hz_wl, 2 Qnt_ac, 3 Tau_ac);
 * returns number of doubles copied, 0 when the state does not exist */
long abo_get_state(const abo_session *s, int which, double *out);

/* ---- building blocks exposed for unit tests (scalar) ---- */
double abo_e_sat(double T);
double abo_q_sat(double T, double p);
double abo_theta_from_z_P0_T_q(double z, double slp, double T, double q);
double abo_rho_air(double T, double q, double p);
double abo_visc_air(double T);
double abo_L_vap(double T);
double abo_cp_air(double q);
double abo_gamma_moist(double T, double q);
double abo_alpha_sw(double T);
double abo_qlw_net(double rlw, double Ts);
double abo_one_on_L(double tha, double qa, double us, double ts, double qs);
double abo_Ri_bulk(double z, double sst, double tha, double ssq, double qa, double ub);
double abo_q_air_rh(double rh, double T, double p);
double abo_q_air_dp(double dp, double p);
double abo_z0tq_LKB(int iflag, double Rer, double z0);
double abo_delta_skin_layer(double alpha, double Qd, double us, int has_qlat, double Qlat);
double abo_psi_m(int algo, double zeta); /* algo id; COARE3P0 and 3P6 share psi */
double abo_psi_h(int algo, double zeta);
double abo_cd_n10_ncar(double w);
double abo_charn_coare3p0(double w);
double abo_charn_coare3p6(double w);
double abo_u_star_andreas(double un10);

/*
 * Direct TURB_* call for one point without skin schemes (what the toy program
 * src/tests/aerobulk_toy.F90 does); out[] = Cd, Ch, Ce, t_zu, q_zu, Ubzu,
 * CdN, ChN, CeN, z0, u*, L, UN10.
 */
void abo_turb_noskin(int algo, int nb_iter, double zt, double zu, double sst, double tha_zt,
                     double ssq, double q_zt, double U_zu, double *out13);

/* Direct TURB_* call on arrays (see .c): t_zt is POTENTIAL temperature, q_zt specific humidity, U_zu the scalar
 * wind, Qsw the NET solar flux; T_s/q_s are in/out; opt[10] = CdN ChN CeN xz0 xu_star xL xUN10 pdT_cs pdT_wl pHz_wl. */
void abo_set_nitend(abo_session *s, int nitend);
int abo_turb(abo_session *s, const char *calgo, int kt, double zt, double zu, long n,
             double *T_s, const double *t_zt, double *q_s, const double *q_zt, const double *U_zu,
             int l_use_cs, int l_use_wl,
             double *Cd, double *Ch, double *Ce, double *t_zu, double *q_zu, double *Ubzu,
             const double *Qsw, const double *rad_lw, const double *slp, int isecday_utc, const double *plong,
             double *const *opt);

/* Station time series (src/tests/test_aerobulk_buoy_series_oce.f90:364-537) for S stations, records [Nt][S];
 * out[28]: rho_zu QL QH Qlw QNS Qsw dT_cs dT_wl TAU dT Hz_wl Qnt_ac Tau_ac Cd Ce Ch theta_zu q_zu t_zu RiB z0 u_star L
 * UN10 Ts Evap q_zt theta_zt (NULL = skip).  hum_kind 0 q, 1 dew-point [K], 2 RH [%]. */
int abo_series(abo_session *s, const char *calgo, int Nt, long S, double zt, double zu,
               const int *isecday_utc, const double *lon,
               const double *sst, const double *t_zt, const double *hum_zt, int hum_kind, const double *wnd,
               const double *slp, const double *rad_sw, const double *rad_lw, int l_skin, double *const *out);

/* ---- sea ice ("parity unpinned": the reference holds no ice fixture, see .c) ---- */
double abo_e_sat_ice(double T);
double abo_q_sat_ice(double T, double p);
double abo_f_m_louis(double zu, double Rib, double Cdn, double z0);
double abo_f_h_louis(double zu, double Rib, double Chn, double z0);
double abo_psi_m_ice(double zeta);
double abo_psi_h_ice(double zeta);
double abo_rough_leng_m(double us, double nua);
int abo_rough_leng_tq(double z0, double us, double nua, double *z0t, double *z0q);
double abo_CdN10_f_LU13(double A);
double abo_CdN_f_LG15_light(double zu, double A, double z0w);
/* TURB_ICE_<nemo|easy|an05|lu12|lg15|lg15_io>; opt[8] = CdN ChN CeN xz0 xu_star xL xUN10 CdN_frm */
int abo_turb_ice(abo_session *s, const char *calgo, double zt, double zu, long n,
                 const double *Ts_i, const double *t_zt, const double *qs_i, const double *q_zt, const double *U_zu,
                 const double *frice, const double *cxn, int per_point_form_drag,
                 double *Cd, double *Ch, double *Ce, double *t_zu, double *q_zu, double *Ubzu, double *const *opt);
/* ice + leads workflow of src/ice/test_aerobulk_oce+ice.f90; out[35], see .c */
int abo_oce_ice(abo_session *s, const char *calgo_ice, const char *calgo_oce, double zt, double zu, long n,
                const double *sit, const double *sst, const double *t_zt, const double *hum_zt, int hum_kind,
                const double *wnd, const double *slp, const double *frice, const double *cxn, int per_point_form_drag,
                double *const *out);

/* sea-ice station series of src/ice/test_aerobulk_buoy_series_ice.f90 on n records; out[21], see .c */
int abo_series_ice(abo_session *s, const char *calgo, double zt, double zu, long n, const double *sic, const double *sit,
                   const double *t_zt, const double *hum_zt, int hum_kind, const double *wnd, const double *slp,
                   const double *rad_sw, const double *rad_lw, double *const *out);

/* test-only: reproduce the pre-drift COARE 3.0 viscosity line (see .c) */
void abo_debug_coare3p0_visc_at_tzu(int on);

#ifdef __cplusplus
}
#endif
#endif
