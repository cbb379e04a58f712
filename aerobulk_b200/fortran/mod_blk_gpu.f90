! mod_blk_gpu.f90 -- drop-in TURB_* routines on the B200 library (SURVEY 8f row 1), for callers such as NEMO's sbcblk
! that call the bulk algorithms directly instead of AEROBULK_MODEL.
!
! SOURCE-ONLY DELIVERABLE (no Fortran compiler in the build image: checked by reading only).
!
! Every routine keeps the dummy-argument list of the reference routine it replaces:
!    TURB_NCAR      src/mod_blk_ncar.f90:57-59          TURB_ANDREAS   src/mod_blk_andreas.f90:66-68
!    TURB_COARE3P0  src/mod_blk_coare3p0.f90:54-59      TURB_COARE3P6  src/mod_blk_coare3p6.f90:123-127
!    TURB_ECMWF     src/mod_blk_ecmwf.f90:63-67
! and forwards to ONE private worker that calls aerobulk_gpu_turb (include/aerobulk_gpu.h).  Arrays must be
! contiguous (whole arrays or contiguous sections), as they are in sbcblk; the worker stops otherwise.
! `nb_iter` and `nitend` are the mod_const globals and are pushed to the library at every call.
MODULE mod_blk_gpu

   USE, INTRINSIC :: iso_c_binding
   USE mod_const,        ONLY: wp, nb_iter, nitend, rdt, gdept_1d
   USE mod_aerobulk_gpu

   IMPLICIT NONE
   PRIVATE

   PUBLIC :: TURB_NCAR, TURB_ANDREAS, TURB_COARE3P0, TURB_COARE3P6, TURB_ECMWF

CONTAINS

   !! pointer of an OPTIONAL contiguous array, C_NULL_PTR when absent
   FUNCTION opt_loc( px ) RESULT( p )
      REAL(wp), DIMENSION(:,:), INTENT(in), OPTIONAL, TARGET :: px
      TYPE(c_ptr) :: p
      p = C_NULL_PTR
      IF( PRESENT(px) ) THEN
         IF( .NOT. IS_CONTIGUOUS(px) ) STOP 'mod_blk_gpu: non-contiguous array argument'
         p = C_LOC(px)
      END IF
   END FUNCTION opt_loc

   SUBROUTINE turb_gpu( calgo, kt, zt, zu, T_s, t_zt, q_s, q_zt, U_zu, l_use_cs, l_use_wl,   &
      &                 Cd, Ch, Ce, t_zu, q_zu, Ubzu,                                        &
      &                 Qsw, rad_lw, slp, pdT_cs, isecday_utc, plong, pdT_wl, pHz_wl,        &
      &                 CdN, ChN, CeN, xz0, xu_star, xL, xUN10 )
      CHARACTER(len=*),         INTENT(in)    :: calgo
      INTEGER,                  INTENT(in)    :: kt
      REAL(wp),                 INTENT(in)    :: zt, zu
      REAL(wp), DIMENSION(:,:), INTENT(inout), TARGET :: T_s, q_s
      REAL(wp), DIMENSION(:,:), INTENT(in),    TARGET :: t_zt, q_zt, U_zu
      LOGICAL,                  INTENT(in)    :: l_use_cs, l_use_wl
      REAL(wp), DIMENSION(:,:), INTENT(out),   TARGET :: Cd, Ch, Ce, t_zu, q_zu, Ubzu
      REAL(wp), DIMENSION(:,:), INTENT(in),    OPTIONAL, TARGET :: Qsw, rad_lw, slp, plong
      INTEGER,                  INTENT(in),    OPTIONAL         :: isecday_utc
      REAL(wp), DIMENSION(:,:), INTENT(out),   OPTIONAL, TARGET :: pdT_cs, pdT_wl, pHz_wl
      REAL(wp), DIMENSION(:,:), INTENT(out),   OPTIONAL, TARGET :: CdN, ChN, CeN, xz0, xu_star, xL, xUN10
      !!
      TYPE(aerobulk_gpu_turb_optional), TARGET :: opt
      CHARACTER(KIND=c_char, LEN=LEN_TRIM(calgo)+1) :: calgo_c
      INTEGER(c_int) :: ierr, ics, iwl, isd
      !!
      calgo_c = TRIM(calgo)//C_NULL_CHAR
      ics = 0_c_int ; IF( l_use_cs ) ics = 1_c_int
      iwl = 0_c_int ; IF( l_use_wl ) iwl = 1_c_int
      isd = 0_c_int ; IF( PRESENT(isecday_utc) ) isd = INT(isecday_utc, c_int)

      !! mod_const globals the library mirrors
      CALL aerobulk_gpu_set_nb_iter( INT(nb_iter, c_int) )
      CALL aerobulk_gpu_set_nitend(  INT(nitend,  c_int) )
      CALL aerobulk_gpu_set_rdt(     REAL(rdt,         c_double) )
      CALL aerobulk_gpu_set_gdept(   REAL(gdept_1d(1), c_double) )

      opt%CdN     = opt_loc(CdN)     ; opt%ChN    = opt_loc(ChN)    ; opt%CeN    = opt_loc(CeN)
      opt%xz0     = opt_loc(xz0)     ; opt%xu_star = opt_loc(xu_star) ; opt%xL   = opt_loc(xL)
      opt%xUN10   = opt_loc(xUN10)
      opt%pdT_cs  = opt_loc(pdT_cs)  ; opt%pdT_wl = opt_loc(pdT_wl) ; opt%pHz_wl = opt_loc(pHz_wl)

      ierr = aerobulk_gpu_turb( calgo_c, INT(kt, c_int), REAL(zt, c_double), REAL(zu, c_double),            &
         &                      INT(SIZE(T_s,1), c_int), INT(SIZE(T_s,2), c_int),                         &
         &                      opt_loc(T_s), opt_loc(t_zt), opt_loc(q_s), opt_loc(q_zt), opt_loc(U_zu),  &
         &                      ics, iwl,                                                                 &
         &                      opt_loc(Cd), opt_loc(Ch), opt_loc(Ce), opt_loc(t_zu), opt_loc(q_zu), opt_loc(Ubzu), &
         &                      opt_loc(Qsw), opt_loc(rad_lw), opt_loc(slp), isd, opt_loc(plong),         &
         &                      C_LOC(opt), 0_c_int )
      !! fail-stop mode (default): the library has already printed the ctl_stop banner and ended the process
      IF( ierr /= 0_c_int ) STOP 'mod_blk_gpu: aerobulk_gpu_turb failed'
   END SUBROUTINE turb_gpu


   SUBROUTINE TURB_NCAR( zt, zu, sst, t_zt, ssq, q_zt, U_zu, Cd, Ch, Ce, t_zu, q_zu, Ubzu,   &
      &                  CdN, ChN, CeN, xz0, xu_star, xL, xUN10 )
      REAL(wp),                 INTENT(in)  :: zt, zu
      REAL(wp), DIMENSION(:,:), INTENT(in)  :: sst, t_zt, ssq, q_zt, U_zu
      REAL(wp), DIMENSION(:,:), INTENT(out) :: Cd, Ch, Ce, t_zu, q_zu, Ubzu
      REAL(wp), DIMENSION(:,:), INTENT(out), OPTIONAL :: CdN, ChN, CeN, xz0, xu_star, xL, xUN10
      REAL(wp), DIMENSION(SIZE(sst,1),SIZE(sst,2)) :: zTs, zqs      ! the worker's in-out pair (unchanged without skin)
      zTs = sst ; zqs = ssq
      CALL turb_gpu( 'ncar', 1, zt, zu, zTs, t_zt, zqs, q_zt, U_zu, .FALSE., .FALSE., Cd, Ch, Ce, t_zu, q_zu, Ubzu,  &
         &           CdN=CdN, ChN=ChN, CeN=CeN, xz0=xz0, xu_star=xu_star, xL=xL, xUN10=xUN10 )
   END SUBROUTINE TURB_NCAR

   SUBROUTINE TURB_ANDREAS( zt, zu, psst, pt_zt, pssq, pq_zt, pU_zu, pCd, pCh, pCe, pt_zu, pq_zu, pUbzu,   &
      &                     pCdN, pChN, pCeN, pz0, pu_star, pL, pUN10 )
      REAL(wp),                 INTENT(in)  :: zt, zu
      REAL(wp), DIMENSION(:,:), INTENT(in)  :: psst, pt_zt, pssq, pq_zt, pU_zu
      REAL(wp), DIMENSION(:,:), INTENT(out) :: pCd, pCh, pCe, pt_zu, pq_zu, pUbzu
      REAL(wp), DIMENSION(:,:), INTENT(out), OPTIONAL :: pCdN, pChN, pCeN, pz0, pu_star, pL, pUN10
      REAL(wp), DIMENSION(SIZE(psst,1),SIZE(psst,2)) :: zTs, zqs
      zTs = psst ; zqs = pssq
      CALL turb_gpu( 'andreas', 1, zt, zu, zTs, pt_zt, zqs, pq_zt, pU_zu, .FALSE., .FALSE.,    &
         &           pCd, pCh, pCe, pt_zu, pq_zu, pUbzu,                                        &
         &           CdN=pCdN, ChN=pChN, CeN=pCeN, xz0=pz0, xu_star=pu_star, xL=pL, xUN10=pUN10 )
   END SUBROUTINE TURB_ANDREAS

   SUBROUTINE TURB_COARE3P6( kt, zt, zu, T_s, t_zt, q_s, q_zt, U_zu, l_use_cs, l_use_wl,   &
      &                      Cd, Ch, Ce, t_zu, q_zu, Ubzu,                                 &
      &                      Qsw, rad_lw, slp, pdT_cs, isecday_utc, plong, pdT_wl, pHz_wl, &
      &                      CdN, ChN, CeN, xz0, xu_star, xL, xUN10 )
      INTEGER,                  INTENT(in)    :: kt
      REAL(wp),                 INTENT(in)    :: zt, zu
      REAL(wp), DIMENSION(:,:), INTENT(inout) :: T_s, q_s
      REAL(wp), DIMENSION(:,:), INTENT(in)    :: t_zt, q_zt, U_zu
      LOGICAL,                  INTENT(in)    :: l_use_cs, l_use_wl
      REAL(wp), DIMENSION(:,:), INTENT(out)   :: Cd, Ch, Ce, t_zu, q_zu, Ubzu
      REAL(wp), DIMENSION(:,:), INTENT(in),  OPTIONAL :: Qsw, rad_lw, slp, plong
      INTEGER,                  INTENT(in),  OPTIONAL :: isecday_utc
      REAL(wp), DIMENSION(:,:), INTENT(out), OPTIONAL :: pdT_cs, pdT_wl, pHz_wl
      REAL(wp), DIMENSION(:,:), INTENT(out), OPTIONAL :: CdN, ChN, CeN, xz0, xu_star, xL, xUN10
      CALL turb_gpu( 'coare3p6', kt, zt, zu, T_s, t_zt, q_s, q_zt, U_zu, l_use_cs, l_use_wl, Cd, Ch, Ce, t_zu, q_zu, Ubzu, &
         &           Qsw=Qsw, rad_lw=rad_lw, slp=slp, pdT_cs=pdT_cs, isecday_utc=isecday_utc, plong=plong,                &
         &           pdT_wl=pdT_wl, pHz_wl=pHz_wl, CdN=CdN, ChN=ChN, CeN=CeN, xz0=xz0, xu_star=xu_star, xL=xL, xUN10=xUN10 )
   END SUBROUTINE TURB_COARE3P6

   SUBROUTINE TURB_COARE3P0( kt, zt, zu, pT_s, pt_zt, pq_s, pq_zt, pU_zu, l_use_cs, l_use_wl,   &
      &                      pCd, pCh, pCe, pt_zu, pq_zu, pUbzu,                                 &
      &                      pQsw, prad_lw, pslp, pdT_cs, isecday_utc, plong, pdT_wl, pHz_wl,    &
      &                      pCdN, pChN, pCeN, pz0, pu_star, pL, pUN10 )
      INTEGER,                  INTENT(in)    :: kt
      REAL(wp),                 INTENT(in)    :: zt, zu
      REAL(wp), DIMENSION(:,:), INTENT(inout) :: pT_s, pq_s
      REAL(wp), DIMENSION(:,:), INTENT(in)    :: pt_zt, pq_zt, pU_zu
      LOGICAL,                  INTENT(in)    :: l_use_cs, l_use_wl
      REAL(wp), DIMENSION(:,:), INTENT(out)   :: pCd, pCh, pCe, pt_zu, pq_zu, pUbzu
      REAL(wp), DIMENSION(:,:), INTENT(in),  OPTIONAL :: pQsw, prad_lw, pslp, plong
      INTEGER,                  INTENT(in),  OPTIONAL :: isecday_utc
      REAL(wp), DIMENSION(:,:), INTENT(out), OPTIONAL :: pdT_cs, pdT_wl, pHz_wl
      REAL(wp), DIMENSION(:,:), INTENT(out), OPTIONAL :: pCdN, pChN, pCeN, pz0, pu_star, pL, pUN10
      CALL turb_gpu( 'coare3p0', kt, zt, zu, pT_s, pt_zt, pq_s, pq_zt, pU_zu, l_use_cs, l_use_wl,                &
         &           pCd, pCh, pCe, pt_zu, pq_zu, pUbzu,                                                          &
         &           Qsw=pQsw, rad_lw=prad_lw, slp=pslp, pdT_cs=pdT_cs, isecday_utc=isecday_utc, plong=plong,     &
         &           pdT_wl=pdT_wl, pHz_wl=pHz_wl, CdN=pCdN, ChN=pChN, CeN=pCeN, xz0=pz0, xu_star=pu_star, xL=pL, xUN10=pUN10 )
   END SUBROUTINE TURB_COARE3P0

   SUBROUTINE TURB_ECMWF( kt, zt, zu, pT_s, pt_zt, pq_s, pq_zt, pU_zu, l_use_cs, l_use_wl,   &
      &                   pCd, pCh, pCe, pt_zu, pq_zu, pUbzu,                                 &
      &                   pQsw, prad_lw, pslp, pdT_cs, pdT_wl, pHz_wl,                        &
      &                   pCdN, pChN, pCeN, pz0, pu_star, pL, pUN10 )
      INTEGER,                  INTENT(in)    :: kt
      REAL(wp),                 INTENT(in)    :: zt, zu
      REAL(wp), DIMENSION(:,:), INTENT(inout) :: pT_s, pq_s
      REAL(wp), DIMENSION(:,:), INTENT(in)    :: pt_zt, pq_zt, pU_zu
      LOGICAL,                  INTENT(in)    :: l_use_cs, l_use_wl
      REAL(wp), DIMENSION(:,:), INTENT(out)   :: pCd, pCh, pCe, pt_zu, pq_zu, pUbzu
      REAL(wp), DIMENSION(:,:), INTENT(in),  OPTIONAL :: pQsw, prad_lw, pslp
      REAL(wp), DIMENSION(:,:), INTENT(out), OPTIONAL :: pdT_cs, pdT_wl, pHz_wl
      REAL(wp), DIMENSION(:,:), INTENT(out), OPTIONAL :: pCdN, pChN, pCeN, pz0, pu_star, pL, pUN10
      CALL turb_gpu( 'ecmwf', kt, zt, zu, pT_s, pt_zt, pq_s, pq_zt, pU_zu, l_use_cs, l_use_wl,          &
         &           pCd, pCh, pCe, pt_zu, pq_zu, pUbzu,                                                 &
         &           Qsw=pQsw, rad_lw=prad_lw, slp=pslp, pdT_cs=pdT_cs, pdT_wl=pdT_wl, pHz_wl=pHz_wl,    &
         &           CdN=pCdN, ChN=pChN, CeN=pCeN, xz0=pz0, xu_star=pu_star, xL=pL, xUN10=pUN10 )
   END SUBROUTINE TURB_ECMWF

END MODULE mod_blk_gpu
