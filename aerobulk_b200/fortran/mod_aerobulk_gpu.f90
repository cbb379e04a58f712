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
      &      aerobulk_gpu_turb, aerobulk_gpu_turb_optional, aerobulk_gpu_set_nitend, &
      &      aerobulk_gpu_set_rdt, aerobulk_gpu_set_gdept, aerobulk_gpu_set_nb_iter, &
      &      aerobulk_gpu_get_nb_iter, aerobulk_gpu_get_use_skin,                 &
      &      aerobulk_gpu_set_device, aerobulk_gpu_set_verbose, aerobulk_gpu_reset,  &
      &      aerobulk_gpu_get_state, aerobulk_gpu_set_state

   !! optional outputs of the TURB_* routines (struct aerobulk_gpu_turb_optional); C_NULL_PTR = not wanted
   TYPE, BIND(C) :: aerobulk_gpu_turb_optional
      TYPE(c_ptr) :: CdN = C_NULL_PTR, ChN = C_NULL_PTR, CeN = C_NULL_PTR
      TYPE(c_ptr) :: xz0 = C_NULL_PTR, xu_star = C_NULL_PTR, xL = C_NULL_PTR, xUN10 = C_NULL_PTR
      TYPE(c_ptr) :: pdT_cs = C_NULL_PTR, pdT_wl = C_NULL_PTR, pHz_wl = C_NULL_PTR
   END TYPE aerobulk_gpu_turb_optional

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
