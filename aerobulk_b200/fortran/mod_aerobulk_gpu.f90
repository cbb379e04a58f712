! mod_aerobulk_gpu.f90 -- ISO_C_BINDING interfaces to libaerobulk_gpu.so (include/aerobulk_gpu.h),
! the B200 (sm_100a) implementation of the AEROBULK_MODEL hot path.
!
! SOURCE-ONLY DELIVERABLE: the build image has no Fortran compiler, so this file has been checked by
! reading only.  It is used by the drop-in `mod_aerobulk.f90` next to it, which keeps the reference's
! AEROBULK_MODEL signature (reference src/mod_aerobulk.f90:176-230) and forwards to these bindings.
!
! Every array argument is passed as TYPE(C_PTR) BY VALUE: C_LOC(array) when present, C_NULL_PTR for an
! absent OPTIONAL -- the C side tests the pointers exactly like the reference tests PRESENT().
MODULE mod_aerobulk_gpu

   USE, INTRINSIC :: iso_c_binding

   IMPLICIT NONE

   PRIVATE

   PUBLIC :: aerobulk_gpu_model, aerobulk_gpu_synchronize,                       &
      &      aerobulk_gpu_init, aerobulk_gpu_bye, aerobulk_gpu_set_devices, aerobulk_gpu_get_devices, aerobulk_gpu_new_session, &
      &      aerobulk_gpu_turb, aerobulk_gpu_turb_optional, aerobulk_gpu_set_nitend, &
      &      aerobulk_gpu_series, aerobulk_gpu_series_out, aerobulk_gpu_series_csv,  &
      &      aerobulk_gpu_turb_ice, aerobulk_gpu_turb_ice_optional,                  &
      &      aerobulk_gpu_oce_ice, aerobulk_gpu_oce_ice_out, aerobulk_gpu_set_ice_form_drag_per_point, &
      &      aerobulk_gpu_series_ice, aerobulk_gpu_series_ice_out,                   &
      &      aerobulk_gpu_set_rdt, aerobulk_gpu_set_gdept, aerobulk_gpu_set_nb_iter, &
      &      aerobulk_gpu_get_nb_iter, aerobulk_gpu_get_use_skin,                 &
      &      aerobulk_gpu_set_device, aerobulk_gpu_set_verbose, aerobulk_gpu_reset,  &
      &      aerobulk_gpu_get_state, aerobulk_gpu_set_state, aerobulk_gpu_host_register, aerobulk_gpu_host_unregister

   !! optional outputs of the TURB_* routines (struct aerobulk_gpu_turb_optional); C_NULL_PTR = not wanted
   TYPE, BIND(C) :: aerobulk_gpu_turb_optional
      TYPE(c_ptr) :: CdN = C_NULL_PTR, ChN = C_NULL_PTR, CeN = C_NULL_PTR
      TYPE(c_ptr) :: xz0 = C_NULL_PTR, xu_star = C_NULL_PTR, xL = C_NULL_PTR, xUN10 = C_NULL_PTR
      TYPE(c_ptr) :: pdT_cs = C_NULL_PTR, pdT_wl = C_NULL_PTR, pHz_wl = C_NULL_PTR
   END TYPE aerobulk_gpu_turb_optional

   !! output series of aerobulk_gpu_series_ice (struct aerobulk_gpu_series_ice_out), n values each; C_NULL_PTR = not wanted
   TYPE, BIND(C) :: aerobulk_gpu_series_ice_out
      TYPE(c_ptr) :: rho_zu = C_NULL_PTR, QL = C_NULL_PTR, QH = C_NULL_PTR, Qlw = C_NULL_PTR, QNS = C_NULL_PTR
      TYPE(c_ptr) :: Qsw = C_NULL_PTR, TAU = C_NULL_PTR, SBLM = C_NULL_PTR
      TYPE(c_ptr) :: Cd_i = C_NULL_PTR, Ch_i = C_NULL_PTR, Ce_i = C_NULL_PTR, z0 = C_NULL_PTR
      TYPE(c_ptr) :: RiB_zt = C_NULL_PTR, RiB_zu = C_NULL_PTR, CdN = C_NULL_PTR
      TYPE(c_ptr) :: u_star = C_NULL_PTR, L = C_NULL_PTR, UN10 = C_NULL_PTR
      TYPE(c_ptr) :: theta_zu = C_NULL_PTR, q_zu = C_NULL_PTR, Ublk = C_NULL_PTR
   END TYPE aerobulk_gpu_series_ice_out

   !! output series of aerobulk_gpu_series (struct aerobulk_gpu_series_out), each [Nt][S]; C_NULL_PTR = not wanted
   TYPE, BIND(C) :: aerobulk_gpu_series_out
      TYPE(c_ptr) :: rho_zu = C_NULL_PTR, QL = C_NULL_PTR, QH = C_NULL_PTR, Qlw = C_NULL_PTR, QNS = C_NULL_PTR
      TYPE(c_ptr) :: Qsw = C_NULL_PTR, dT_cs = C_NULL_PTR, dT_wl = C_NULL_PTR, TAU = C_NULL_PTR, dT = C_NULL_PTR
      TYPE(c_ptr) :: Hz_wl = C_NULL_PTR, Qnt_ac = C_NULL_PTR, Tau_ac = C_NULL_PTR
      TYPE(c_ptr) :: Cd = C_NULL_PTR, Ce = C_NULL_PTR, Ch = C_NULL_PTR
      TYPE(c_ptr) :: theta_zu = C_NULL_PTR, q_zu = C_NULL_PTR, t_zu = C_NULL_PTR, RiB = C_NULL_PTR, z0 = C_NULL_PTR
      TYPE(c_ptr) :: u_star = C_NULL_PTR, L = C_NULL_PTR, UN10 = C_NULL_PTR, Ts = C_NULL_PTR, Evap = C_NULL_PTR
      TYPE(c_ptr) :: q_zt = C_NULL_PTR, theta_zt = C_NULL_PTR
   END TYPE aerobulk_gpu_series_out

   !! optional outputs of the TURB_ICE_* routines (struct aerobulk_gpu_turb_ice_optional)
   TYPE, BIND(C) :: aerobulk_gpu_turb_ice_optional
      TYPE(c_ptr) :: CdN = C_NULL_PTR, ChN = C_NULL_PTR, CeN = C_NULL_PTR
      TYPE(c_ptr) :: xz0 = C_NULL_PTR, xu_star = C_NULL_PTR, xL = C_NULL_PTR, xUN10 = C_NULL_PTR, CdN_frm = C_NULL_PTR
   END TYPE aerobulk_gpu_turb_ice_optional

   !! outputs of aerobulk_gpu_oce_ice (struct aerobulk_gpu_oce_ice_out): _i over the ice, _w over the leads, cell means
   TYPE, BIND(C) :: aerobulk_gpu_oce_ice_out
      TYPE(c_ptr) :: Cd_i = C_NULL_PTR, Ch_i = C_NULL_PTR, Ce_i = C_NULL_PTR, theta_zu_i = C_NULL_PTR, q_zu_i = C_NULL_PTR
      TYPE(c_ptr) :: t_zu_i = C_NULL_PTR, Ub_i = C_NULL_PTR, RiB_i = C_NULL_PTR, z0_i = C_NULL_PTR, u_star_i = C_NULL_PTR
      TYPE(c_ptr) :: L_i = C_NULL_PTR, UN10_i = C_NULL_PTR, rho_zu_i = C_NULL_PTR
      TYPE(c_ptr) :: Tau_i = C_NULL_PTR, QH_i = C_NULL_PTR, QL_i = C_NULL_PTR, Evap_i = C_NULL_PTR
      TYPE(c_ptr) :: Cd_w = C_NULL_PTR, Ch_w = C_NULL_PTR, Ce_w = C_NULL_PTR, theta_zu_w = C_NULL_PTR, q_zu_w = C_NULL_PTR
      TYPE(c_ptr) :: Ub_w = C_NULL_PTR, z0_w = C_NULL_PTR, u_star_w = C_NULL_PTR, L_w = C_NULL_PTR, UN10_w = C_NULL_PTR
      TYPE(c_ptr) :: Tau_w = C_NULL_PTR, QH_w = C_NULL_PTR, QL_w = C_NULL_PTR, Evap_w = C_NULL_PTR
      TYPE(c_ptr) :: Tau = C_NULL_PTR, QH = C_NULL_PTR, QL = C_NULL_PTR, Evap = C_NULL_PTR
   END TYPE aerobulk_gpu_oce_ice_out

   INTERFACE

      !! int aerobulk_gpu_model(int jt, int Nt, const char *calgo, double zt, double zu, int Ni, int Nj,
      !!                        6 x const double*, 5 x double*, const int *Niter, const int *l_use_skin,
      !!                        const double *rad_sw, const double *rad_lw, double *T_s)
      FUNCTION aerobulk_gpu_model( jt, Nt, calgo, zt, zu, Ni, Nj,           &
         &                         sst, t_zt, hum_zt, U_zu, V_zu, slp,      &
         &                         QL, QH, Tau_x, Tau_y, Evap,              &
         &                         Niter, l_use_skin, rad_sw, rad_lw, T_s ) &
         &     BIND(C, NAME='aerobulk_gpu_model') RESULT(ierr)
         IMPORT :: c_int, c_double, c_char, c_ptr
         INTEGER(c_int),         VALUE                     :: jt, Nt
         CHARACTER(KIND=c_char), DIMENSION(*), INTENT(in)  :: calgo      !: NUL-terminated
         REAL(c_double),         VALUE                     :: zt, zu
         INTEGER(c_int),         VALUE                     :: Ni, Nj
         TYPE(c_ptr),            VALUE                     :: sst, t_zt, hum_zt, U_zu, V_zu, slp
         TYPE(c_ptr),            VALUE                     :: QL, QH, Tau_x, Tau_y, Evap
         TYPE(c_ptr),            VALUE                     :: Niter, l_use_skin      !: int*, NULL if absent
         TYPE(c_ptr),            VALUE                     :: rad_sw, rad_lw, T_s    !: double*, NULL if absent
         INTEGER(c_int)                                    :: ierr
      END FUNCTION aerobulk_gpu_model

      !! int aerobulk_gpu_init(int Nt, const char *calgo, int Ni, int Nj, 6 x const double*, const int *l_use_skin,
      !!                       const double *rad_sw, const double *rad_lw)   -- AEROBULK_INIT on host arrays
      FUNCTION aerobulk_gpu_init( Nt, calgo, Ni, Nj, sst, t_zt, hum_zt, U_zu, V_zu, slp, l_use_skin, rad_sw, rad_lw ) &
         &     BIND(C, NAME='aerobulk_gpu_init') RESULT(ierr)
         IMPORT :: c_int, c_char, c_ptr
         INTEGER(c_int),         VALUE                     :: Nt
         CHARACTER(KIND=c_char), DIMENSION(*), INTENT(in)  :: calgo
         INTEGER(c_int),         VALUE                     :: Ni, Nj
         TYPE(c_ptr),            VALUE                     :: sst, t_zt, hum_zt, U_zu, V_zu, slp
         TYPE(c_ptr),            VALUE                     :: l_use_skin, rad_sw, rad_lw
         INTEGER(c_int)                                    :: ierr
      END FUNCTION aerobulk_gpu_init

      SUBROUTINE aerobulk_gpu_bye() BIND(C, NAME='aerobulk_gpu_bye')
      END SUBROUTINE aerobulk_gpu_bye

      !! Split every AEROBULK_MODEL call over n GPUs inside the library (row blocks; include/aerobulk_gpu.h)
      FUNCTION aerobulk_gpu_set_devices( n ) BIND(C, NAME='aerobulk_gpu_set_devices') RESULT(ierr)
         IMPORT :: c_int
         INTEGER(c_int), VALUE :: n
         INTEGER(c_int)        :: ierr
      END FUNCTION aerobulk_gpu_set_devices

      FUNCTION aerobulk_gpu_get_devices() BIND(C, NAME='aerobulk_gpu_get_devices') RESULT(n)
         IMPORT :: c_int
         INTEGER(c_int) :: n
      END FUNCTION aerobulk_gpu_get_devices

      SUBROUTINE aerobulk_gpu_new_session() BIND(C, NAME='aerobulk_gpu_new_session')
      END SUBROUTINE aerobulk_gpu_new_session

      !! Direct TURB_COARE3P0 / TURB_COARE3P6 / TURB_ECMWF / TURB_NCAR / TURB_ANDREAS (selected by calgo);
      !! pt_zt: POTENTIAL temperature, pQsw: NET solar flux, pT_s/pq_s in-out; popt: C_LOC of a
      !! TYPE(aerobulk_gpu_turb_optional) or C_NULL_PTR; on_device = 0 for host arrays.
      FUNCTION aerobulk_gpu_turb( calgo, kt, zt, zu, Ni, Nj, pT_s, pt_zt, pq_s, pq_zt, pU_zu, l_use_cs, l_use_wl, &
         &                        pCd, pCh, pCe, pt_zu, pq_zu, pUbzu, pQsw, prad_lw, pslp, isecday_utc, plong,    &
         &                        popt, on_device ) BIND(C, NAME='aerobulk_gpu_turb') RESULT(ierr)
         IMPORT :: c_int, c_double, c_char, c_ptr
         CHARACTER(KIND=c_char), DIMENSION(*), INTENT(in) :: calgo
         INTEGER(c_int), VALUE :: kt
         REAL(c_double), VALUE :: zt, zu
         INTEGER(c_int), VALUE :: Ni, Nj
         TYPE(c_ptr),    VALUE :: pT_s, pt_zt, pq_s, pq_zt, pU_zu
         INTEGER(c_int), VALUE :: l_use_cs, l_use_wl
         TYPE(c_ptr),    VALUE :: pCd, pCh, pCe, pt_zu, pq_zu, pUbzu
         TYPE(c_ptr),    VALUE :: pQsw, prad_lw, pslp
         INTEGER(c_int), VALUE :: isecday_utc
         TYPE(c_ptr),    VALUE :: plong, popt
         INTEGER(c_int), VALUE :: on_device
         INTEGER(c_int)        :: ierr
      END FUNCTION aerobulk_gpu_turb

      !! Station time series (time loop of src/tests/test_aerobulk_buoy_series_oce.f90:364-537 for S stations);
      !! records are [Nt][S], i.e. Fortran arrays dimensioned (S,Nt); pisd: C_LOC of INTEGER(c_int) isecday_utc(Nt);
      !! pout: C_LOC of a TYPE(aerobulk_gpu_series_out); hum_kind 0 q / 1 dew-point [K] / 2 RH [%].
      FUNCTION aerobulk_gpu_series( calgo, Nt, S, zt, zu, pisd, plon, psst, pt_zt, phum_zt, hum_kind, pwind, pslp,   &
         &                          prad_sw, prad_lw, l_use_skin, pout, on_device )                                  &
         &     BIND(C, NAME='aerobulk_gpu_series') RESULT(ierr)
         IMPORT :: c_int, c_long_long, c_double, c_char, c_ptr
         CHARACTER(KIND=c_char), DIMENSION(*), INTENT(in) :: calgo
         INTEGER(c_int),       VALUE :: Nt
         INTEGER(c_long_long), VALUE :: S
         REAL(c_double),       VALUE :: zt, zu
         TYPE(c_ptr),          VALUE :: pisd, plon, psst, pt_zt, phum_zt
         INTEGER(c_int),       VALUE :: hum_kind
         TYPE(c_ptr),          VALUE :: pwind, pslp, prad_sw, prad_lw
         INTEGER(c_int),       VALUE :: l_use_skin
         TYPE(c_ptr),          VALUE :: pout
         INTEGER(c_int),       VALUE :: on_device
         INTEGER(c_int)              :: ierr
      END FUNCTION aerobulk_gpu_series

      !! One station, CSV in / CSV out (the buoy-series program with text files for NetCDF)
      FUNCTION aerobulk_gpu_series_csv( path_in, path_out, calgo, zt, zu, l_use_skin ) &
         &     BIND(C, NAME='aerobulk_gpu_series_csv') RESULT(ierr)
         IMPORT :: c_int, c_double, c_char
         CHARACTER(KIND=c_char), DIMENSION(*), INTENT(in) :: path_in, path_out, calgo   !: NUL-terminated
         REAL(c_double), VALUE :: zt, zu
         INTEGER(c_int), VALUE :: l_use_skin
         INTEGER(c_int)        :: ierr
      END FUNCTION aerobulk_gpu_series_csv

      !! TURB_ICE_NEMO / _EASY / _AN05 / _LU12 / _LG15 / _LG15_IO (src/ice/mod_blk_ice_*.f90), selected by calgo;
      !! pfrice: lu12 / lg15 only, pCxN_easy: C_LOC of REAL(c_double) (/CdN,ChN,CeN/) for easy; C_NULL_PTR otherwise
      FUNCTION aerobulk_gpu_turb_ice( calgo, zt, zu, Ni, Nj, pTs_i, pt_zt, pqs_i, pq_zt, pU_zu, pfrice, pCxN_easy,   &
         &                            pCd, pCh, pCe, pt_zu, pq_zu, pUbzu, popt, on_device )                          &
         &     BIND(C, NAME='aerobulk_gpu_turb_ice') RESULT(ierr)
         IMPORT :: c_int, c_double, c_char, c_ptr
         CHARACTER(KIND=c_char), DIMENSION(*), INTENT(in) :: calgo
         REAL(c_double), VALUE :: zt, zu
         INTEGER(c_int), VALUE :: Ni, Nj
         TYPE(c_ptr),    VALUE :: pTs_i, pt_zt, pqs_i, pq_zt, pU_zu, pfrice, pCxN_easy
         TYPE(c_ptr),    VALUE :: pCd, pCh, pCe, pt_zu, pq_zu, pUbzu, popt
         INTEGER(c_int), VALUE :: on_device
         INTEGER(c_int)        :: ierr
      END FUNCTION aerobulk_gpu_turb_ice

      !! Ice + leads fluxes (src/ice/test_aerobulk_oce+ice.f90:225-412) on n points; calgo_oce: C_NULL_CHAR-terminated
      !! ocean algorithm for the leads; pout: C_LOC of a TYPE(aerobulk_gpu_oce_ice_out)
      FUNCTION aerobulk_gpu_oce_ice( calgo_ice, calgo_oce, zt, zu, n, psit, psst, pt_zt, phum_zt, hum_kind, pwind,   &
         &                           pslp, pfrice, pCxN_easy, pout, on_device )                                      &
         &     BIND(C, NAME='aerobulk_gpu_oce_ice') RESULT(ierr)
         IMPORT :: c_int, c_long_long, c_double, c_char, c_ptr
         CHARACTER(KIND=c_char), DIMENSION(*), INTENT(in) :: calgo_ice, calgo_oce
         REAL(c_double),       VALUE :: zt, zu
         INTEGER(c_long_long), VALUE :: n
         TYPE(c_ptr),          VALUE :: psit, psst, pt_zt, phum_zt
         INTEGER(c_int),       VALUE :: hum_kind
         TYPE(c_ptr),          VALUE :: pwind, pslp, pfrice, pCxN_easy, pout
         INTEGER(c_int),       VALUE :: on_device
         INTEGER(c_int)              :: ierr
      END FUNCTION aerobulk_gpu_oce_ice

      SUBROUTINE aerobulk_gpu_set_ice_form_drag_per_point( on ) BIND(C, NAME='aerobulk_gpu_set_ice_form_drag_per_point')
         IMPORT :: c_int
         INTEGER(c_int), VALUE :: on
      END SUBROUTINE aerobulk_gpu_set_ice_form_drag_per_point

      !! sea-ice station series (src/ice/test_aerobulk_buoy_series_ice.f90) on n records; hum_kind 0 q, 1 dew-point, 2 RH
      FUNCTION aerobulk_gpu_series_ice( calgo, zt, zu, n, psic, psit, pt_zt, phum_zt, hum_kind, pwind, pslp,  &
         &                              prad_sw, prad_lw, pout, on_device )                                  &
         &     BIND(C, NAME='aerobulk_gpu_series_ice') RESULT(ierr)
         IMPORT :: c_int, c_long_long, c_double, c_char, c_ptr
         CHARACTER(KIND=c_char), DIMENSION(*), INTENT(in) :: calgo      !! NUL-terminated
         REAL(c_double),       VALUE :: zt, zu
         INTEGER(c_long_long), VALUE :: n
         TYPE(c_ptr),          VALUE :: psic, psit, pt_zt, phum_zt      !! C_LOC of the (host or device) arrays
         INTEGER(c_int),       VALUE :: hum_kind
         TYPE(c_ptr),          VALUE :: pwind, pslp, prad_sw, prad_lw
         TYPE(c_ptr),          VALUE :: pout                            !! C_LOC of a TYPE(aerobulk_gpu_series_ice_out)
         INTEGER(c_int),       VALUE :: on_device
         INTEGER(c_int)              :: ierr
      END FUNCTION aerobulk_gpu_series_ice

      !! page-lock / release an existing array: ierr = aerobulk_gpu_host_register( C_LOC(sst), INT(8*SIZE(sst), c_size_t) )
      FUNCTION aerobulk_gpu_host_register( ptr, nbytes ) BIND(C, NAME='aerobulk_gpu_host_register') RESULT(ierr)
         IMPORT :: c_int, c_ptr, c_size_t
         TYPE(c_ptr),       VALUE :: ptr
         INTEGER(c_size_t), VALUE :: nbytes
         INTEGER(c_int)           :: ierr
      END FUNCTION aerobulk_gpu_host_register
      FUNCTION aerobulk_gpu_host_unregister( ptr ) BIND(C, NAME='aerobulk_gpu_host_unregister') RESULT(ierr)
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: ptr
         INTEGER(c_int)     :: ierr
      END FUNCTION aerobulk_gpu_host_unregister

      SUBROUTINE aerobulk_gpu_set_nitend( knitend ) BIND(C, NAME='aerobulk_gpu_set_nitend')
         IMPORT :: c_int
         INTEGER(c_int), VALUE :: knitend
      END SUBROUTINE aerobulk_gpu_set_nitend

      FUNCTION aerobulk_gpu_synchronize() BIND(C, NAME='aerobulk_gpu_synchronize') RESULT(ierr)
         IMPORT :: c_int
         INTEGER(c_int) :: ierr
      END FUNCTION aerobulk_gpu_synchronize

      SUBROUTINE aerobulk_gpu_set_rdt( rdt ) BIND(C, NAME='aerobulk_gpu_set_rdt')
         IMPORT :: c_double
         REAL(c_double), VALUE :: rdt
      END SUBROUTINE aerobulk_gpu_set_rdt

      SUBROUTINE aerobulk_gpu_set_gdept( gdept ) BIND(C, NAME='aerobulk_gpu_set_gdept')
         IMPORT :: c_double
         REAL(c_double), VALUE :: gdept
      END SUBROUTINE aerobulk_gpu_set_gdept

      SUBROUTINE aerobulk_gpu_set_nb_iter( kiter ) BIND(C, NAME='aerobulk_gpu_set_nb_iter')
         IMPORT :: c_int
         INTEGER(c_int), VALUE :: kiter
      END SUBROUTINE aerobulk_gpu_set_nb_iter

      FUNCTION aerobulk_gpu_get_nb_iter() BIND(C, NAME='aerobulk_gpu_get_nb_iter') RESULT(kiter)
         IMPORT :: c_int
         INTEGER(c_int) :: kiter
      END FUNCTION aerobulk_gpu_get_nb_iter

      FUNCTION aerobulk_gpu_get_use_skin() BIND(C, NAME='aerobulk_gpu_get_use_skin') RESULT(kskin)
         IMPORT :: c_int
         INTEGER(c_int) :: kskin
      END FUNCTION aerobulk_gpu_get_use_skin

      FUNCTION aerobulk_gpu_set_device( kdev ) BIND(C, NAME='aerobulk_gpu_set_device') RESULT(ierr)
         IMPORT :: c_int
         INTEGER(c_int), VALUE :: kdev
         INTEGER(c_int)        :: ierr
      END FUNCTION aerobulk_gpu_set_device

      SUBROUTINE aerobulk_gpu_set_verbose( kon ) BIND(C, NAME='aerobulk_gpu_set_verbose')
         IMPORT :: c_int
         INTEGER(c_int), VALUE :: kon
      END SUBROUTINE aerobulk_gpu_set_verbose

      SUBROUTINE aerobulk_gpu_reset() BIND(C, NAME='aerobulk_gpu_reset')
      END SUBROUTINE aerobulk_gpu_reset

      !! long aerobulk_gpu_get_state(int which, double *host_out, long n)   which: 0 dT_wl 1 Hz_wl 2 Qnt_ac 3 Tau_ac
      FUNCTION aerobulk_gpu_get_state( kwhich, pout, kn ) BIND(C, NAME='aerobulk_gpu_get_state') RESULT(kcopied)
         IMPORT :: c_int, c_long, c_ptr
         INTEGER(c_int),  VALUE :: kwhich
         TYPE(c_ptr),     VALUE :: pout
         INTEGER(c_long), VALUE :: kn
         INTEGER(c_long)        :: kcopied
      END FUNCTION aerobulk_gpu_get_state

      FUNCTION aerobulk_gpu_set_state( kwhich, pin, kn ) BIND(C, NAME='aerobulk_gpu_set_state') RESULT(kcopied)
         IMPORT :: c_int, c_long, c_ptr
         INTEGER(c_int),  VALUE :: kwhich
         TYPE(c_ptr),     VALUE :: pin
         INTEGER(c_long), VALUE :: kn
         INTEGER(c_long)        :: kcopied
      END FUNCTION aerobulk_gpu_set_state

   END INTERFACE

END MODULE mod_aerobulk_gpu
