! mod_aerobulk.f90 -- DROP-IN replacement of the reference's src/mod_aerobulk.f90: same module name,
! same public entry points, same AEROBULK_MODEL signature (reference src/mod_aerobulk.f90:176-230);
! the body forwards to the CUDA implementation through mod_aerobulk_gpu.f90.
!
! SOURCE-ONLY DELIVERABLE (no Fortran compiler in the build image; checked by reading only).
!
! How it maps onto the reference:
!   * AEROBULK_INIT's work (skin-scheme decision, mask, humidity-type detection, unit checks,
!     src/mod_aerobulk.f90:24-160) is done by the library at jt==1, on the device, with the same
!     fail-stop messages; the globals it sets in mod_const (nb_iter, nitend, l_use_skin_schemes,
!     rdt, gdept_1d) stay authoritative: they are pushed to / mirrored from the library at each call.
!   * aerobulk_compute + TURB_* + BULK_FORMULA (src/mod_aerobulk_compute.f90:22-213) are one fused
!     kernel launch per call; the warm-layer state lives on the device between jt==1 and jt==Nt.
!   * Arrays are passed without copy when they are contiguous (IS_CONTIGUOUS); array sections are
!     copied in / out through contiguous temporaries.
!
! Build (replacing libaerobulk.a's mod_aerobulk.o), e.g.:
!   gfortran -O2 -fdefault-real-8 -c mod_const.f90 mod_aerobulk_gpu.f90 mod_aerobulk.f90
!   gfortran my_prog.f90 mod_aerobulk.o mod_aerobulk_gpu.o mod_const.o -L<repo>/aerobulk_b200 -laerobulk_gpu
MODULE mod_aerobulk

   USE, INTRINSIC :: iso_c_binding
   USE mod_const        !: wp, nb_iter, nitend, rdt, gdept_1d, l_use_skin_schemes (reference module, unchanged)
   USE mod_aerobulk_gpu

   IMPLICIT NONE

   PRIVATE

   PUBLIC :: AEROBULK_INIT, AEROBULK_MODEL, AEROBULK_BYE

CONTAINS

   SUBROUTINE AEROBULK_INIT( Nt, calgo, psst, pta, pha, pU, pV, pslp,  l_use_skin, prsw, prlw )
      !! Same interface and behaviour as the reference (src/mod_aerobulk.f90:24-160): skin flag and its two errors,
      !! nitend = Nt, sanity mask, humidity type, unit checks -- done by the library on the GPU (aerobulk_gpu_init),
      !! with the reference's messages and fail-stop.  As in the reference, AEROBULK_MODEL(jt==1) runs the same
      !! initialisation again on its own arguments, so calling this routine first is never required.
      INTEGER,                  INTENT(in)  :: Nt
      CHARACTER(len=*),         INTENT(in)  :: calgo
      REAL(wp), DIMENSION(:,:), INTENT(in), TARGET :: psst, pta, pha, pU, pV, pslp
      LOGICAL,                  INTENT(in), OPTIONAL :: l_use_skin
      REAL(wp), DIMENSION(:,:), INTENT(in), OPTIONAL, TARGET :: prsw, prlw
      !!
      INTEGER :: Ni, Nj, ierr
      INTEGER(c_int), TARGET :: iskin
      TYPE(c_ptr) :: pskin, p_rsw, p_rlw
      CHARACTER(KIND=c_char, LEN=LEN_TRIM(calgo)+1) :: calgo_c
      !! contiguous copies (AEROBULK_INIT is called once per run: always copied, no IS_CONTIGUOUS case analysis)
      REAL(wp), DIMENSION(:,:), ALLOCATABLE, TARGET :: c_sst, c_ta, c_ha, c_U, c_V, c_slp, c_rsw, c_rlw
      Ni = SIZE(psst,1)
      Nj = SIZE(psst,2)
      IF( ANY( (/ SIZE(pta,1),SIZE(pha,1),SIZE(pU,1),SIZE(pV,1),SIZE(pslp,1) /) /= Ni ) .OR. &
         &ANY( (/ SIZE(pta,2),SIZE(pha,2),SIZE(pU,2),SIZE(pV,2),SIZE(pslp,2) /) /= Nj ) )    &
         &  STOP 'AEROBULK_INIT => SST, t_air, hum, U, V and SLP arrays do not agree in shape!'   ! reference :85-92
      ALLOCATE( c_sst(Ni,Nj), c_ta(Ni,Nj), c_ha(Ni,Nj), c_U(Ni,Nj), c_V(Ni,Nj), c_slp(Ni,Nj) )
      c_sst = psst ; c_ta = pta ; c_ha = pha ; c_U = pU ; c_V = pV ; c_slp = pslp
      pskin = C_NULL_PTR
      IF( PRESENT(l_use_skin) ) THEN
         iskin = 0_c_int
         IF( l_use_skin ) iskin = 1_c_int
         pskin = C_LOC(iskin)
      END IF
      p_rsw = C_NULL_PTR
      p_rlw = C_NULL_PTR
      IF( PRESENT(prsw) .AND. PRESENT(prlw) ) THEN
         ALLOCATE( c_rsw(Ni,Nj), c_rlw(Ni,Nj) )
         c_rsw = prsw ; c_rlw = prlw
         p_rsw = C_LOC(c_rsw)
         p_rlw = C_LOC(c_rlw)
      END IF
      calgo_c = TRIM(calgo)//C_NULL_CHAR
      ierr = aerobulk_gpu_init( INT(Nt,c_int), calgo_c, INT(Ni,c_int), INT(Nj,c_int), C_LOC(c_sst), C_LOC(c_ta),       &
         &                      C_LOC(c_ha), C_LOC(c_U), C_LOC(c_V), C_LOC(c_slp), pskin, p_rsw, p_rlw )
      IF( ierr /= 0 ) STOP 'AEROBULK_INIT (GPU): the library reported an error'
      !! mirror the session globals the reference keeps in mod_const
      nitend = Nt
      l_use_skin_schemes = ( aerobulk_gpu_get_use_skin() /= 0_c_int )
   END SUBROUTINE AEROBULK_INIT


   SUBROUTINE AEROBULK_BYE()
      !! reference :162-170 (AEROBULK_MODEL prints the same banner itself at jt==Nt)
      CALL aerobulk_gpu_bye()
   END SUBROUTINE AEROBULK_BYE


   SUBROUTINE AEROBULK_MODEL( jt, Nt, &
      &                       calgo, zt, zu, sst, t_zt,   &
      &                       hum_zt, U_zu, V_zu, slp,    &
      &                       QL, QH, Tau_x, Tau_y, Evap, &
      &                       Niter, l_use_skin, rad_sw, rad_lw, T_s  )
      !!======================================================================================
      !! Same arguments, units and OPTIONAL semantics as the reference (src/mod_aerobulk.f90:181-230):
      !!  jt,Nt : current / total number of time records        calgo : 'coare3p0' 'coare3p6' 'ncar' 'ecmwf' 'andreas'
      !!  zt,zu : measurement heights [m]                      sst, t_zt [K]   hum_zt [kg/kg | % | K]
      !!  U_zu,V_zu [m/s]  slp [Pa]                            QL,QH [W/m^2]   Tau_x,Tau_y [N/m^2]  Evap [kg/m^2/s]
      !!  Niter : iterations (sticky)   l_use_skin + rad_sw + rad_lw [W/m^2] => skin schemes, T_s [K] out
      !!======================================================================================
      INTEGER,                  INTENT(in)  :: jt, Nt
      CHARACTER(len=*),         INTENT(in)  :: calgo
      REAL(wp),                 INTENT(in)  :: zt, zu
      REAL(wp), DIMENSION(:,:), INTENT(in),  TARGET :: sst, t_zt, hum_zt, U_zu, V_zu, slp
      REAL(wp), DIMENSION(:,:), INTENT(out), TARGET :: QL, QH, Tau_x, Tau_y, Evap
      INTEGER,                  INTENT(in),  OPTIONAL :: Niter
      LOGICAL,                  INTENT(in),  OPTIONAL :: l_use_skin
      REAL(wp), DIMENSION(:,:), INTENT(in),  OPTIONAL, TARGET :: rad_sw, rad_lw
      REAL(wp), DIMENSION(:,:), INTENT(out), OPTIONAL, TARGET :: T_s
      !!
      INTEGER :: Ni, Nj, ierr
      INTEGER(c_int), TARGET :: iNiter, iskin
      TYPE(c_ptr) :: pNiter, pskin, prsw, prlw, pTs
      TYPE(c_ptr) :: pin(6), pout(5)
      CHARACTER(KIND=c_char, LEN=LEN_TRIM(calgo)+1) :: calgo_c
      !! contiguous temporaries, only allocated for non-contiguous actual arguments:
      REAL(wp), DIMENSION(:,:), ALLOCATABLE, TARGET :: c_sst, c_tzt, c_hum, c_U, c_V, c_slp, c_rsw, c_rlw
      REAL(wp), DIMENSION(:,:), ALLOCATABLE, TARGET :: c_QL, c_QH, c_Tx, c_Ty, c_Ev, c_Ts
      LOGICAL :: lsrad
      !!======================================================================================
      Ni = SIZE(sst,1)
      Nj = SIZE(sst,2)

      IF( PRESENT(Niter) ) nb_iter = Niter     ! sticky global of mod_const, as in the reference (:236)
      iNiter = INT(nb_iter, c_int)
      pNiter = C_LOC(iNiter)

      pskin = C_NULL_PTR
      IF( PRESENT(l_use_skin) ) THEN
         iskin = 0_c_int
         IF( l_use_skin ) iskin = 1_c_int
         pskin = C_LOC(iskin)
      END IF

      !! module globals that callers may have overwritten in mod_const (reference :31-33):
      CALL aerobulk_gpu_set_rdt(   REAL(rdt,         c_double) )
      CALL aerobulk_gpu_set_gdept( REAL(gdept_1d(1), c_double) )

      calgo_c = TRIM(calgo)//C_NULL_CHAR

      !! ---- inputs
      IF( IS_CONTIGUOUS(sst) )    THEN ; pin(1) = C_LOC(sst)
      ELSE ; ALLOCATE(c_sst(Ni,Nj)) ; c_sst = sst    ; pin(1) = C_LOC(c_sst) ; END IF
      IF( IS_CONTIGUOUS(t_zt) )   THEN ; pin(2) = C_LOC(t_zt)
      ELSE ; ALLOCATE(c_tzt(Ni,Nj)) ; c_tzt = t_zt   ; pin(2) = C_LOC(c_tzt) ; END IF
      IF( IS_CONTIGUOUS(hum_zt) ) THEN ; pin(3) = C_LOC(hum_zt)
      ELSE ; ALLOCATE(c_hum(Ni,Nj)) ; c_hum = hum_zt ; pin(3) = C_LOC(c_hum) ; END IF
      IF( IS_CONTIGUOUS(U_zu) )   THEN ; pin(4) = C_LOC(U_zu)
      ELSE ; ALLOCATE(c_U(Ni,Nj))   ; c_U = U_zu     ; pin(4) = C_LOC(c_U)   ; END IF
      IF( IS_CONTIGUOUS(V_zu) )   THEN ; pin(5) = C_LOC(V_zu)
      ELSE ; ALLOCATE(c_V(Ni,Nj))   ; c_V = V_zu     ; pin(5) = C_LOC(c_V)   ; END IF
      IF( IS_CONTIGUOUS(slp) )    THEN ; pin(6) = C_LOC(slp)
      ELSE ; ALLOCATE(c_slp(Ni,Nj)) ; c_slp = slp    ; pin(6) = C_LOC(c_slp) ; END IF

      lsrad = ( PRESENT(rad_sw) .AND. PRESENT(rad_lw) )   ! reference :242
      prsw = C_NULL_PTR
      prlw = C_NULL_PTR
      pTs  = C_NULL_PTR
      IF( lsrad ) THEN
         IF( IS_CONTIGUOUS(rad_sw) ) THEN ; prsw = C_LOC(rad_sw)
         ELSE ; ALLOCATE(c_rsw(Ni,Nj)) ; c_rsw = rad_sw ; prsw = C_LOC(c_rsw) ; END IF
         IF( IS_CONTIGUOUS(rad_lw) ) THEN ; prlw = C_LOC(rad_lw)
         ELSE ; ALLOCATE(c_rlw(Ni,Nj)) ; c_rlw = rad_lw ; prlw = C_LOC(c_rlw) ; END IF
         IF( PRESENT(T_s) ) THEN
            IF( IS_CONTIGUOUS(T_s) ) THEN ; pTs = C_LOC(T_s)
            ELSE ; ALLOCATE(c_Ts(Ni,Nj)) ; pTs = C_LOC(c_Ts) ; END IF
         END IF
      END IF

      !! ---- outputs
      IF( IS_CONTIGUOUS(QL) )    THEN ; pout(1) = C_LOC(QL)
      ELSE ; ALLOCATE(c_QL(Ni,Nj)) ; pout(1) = C_LOC(c_QL) ; END IF
      IF( IS_CONTIGUOUS(QH) )    THEN ; pout(2) = C_LOC(QH)
      ELSE ; ALLOCATE(c_QH(Ni,Nj)) ; pout(2) = C_LOC(c_QH) ; END IF
      IF( IS_CONTIGUOUS(Tau_x) ) THEN ; pout(3) = C_LOC(Tau_x)
      ELSE ; ALLOCATE(c_Tx(Ni,Nj)) ; pout(3) = C_LOC(c_Tx) ; END IF
      IF( IS_CONTIGUOUS(Tau_y) ) THEN ; pout(4) = C_LOC(Tau_y)
      ELSE ; ALLOCATE(c_Ty(Ni,Nj)) ; pout(4) = C_LOC(c_Ty) ; END IF
      IF( IS_CONTIGUOUS(Evap) )  THEN ; pout(5) = C_LOC(Evap)
      ELSE ; ALLOCATE(c_Ev(Ni,Nj)) ; pout(5) = C_LOC(c_Ev) ; END IF

      !! ---- the GPU call (fail-stop inside the library, like ctl_stop / STOP in the reference)
      ierr = aerobulk_gpu_model( INT(jt,c_int), INT(Nt,c_int), calgo_c, REAL(zt,c_double), REAL(zu,c_double), &
         &                       INT(Ni,c_int), INT(Nj,c_int),                                               &
         &                       pin(1), pin(2), pin(3), pin(4), pin(5), pin(6),                             &
         &                       pout(1), pout(2), pout(3), pout(4), pout(5),                                &
         &                       pNiter, pskin, prsw, prlw, pTs )
      IF( ierr /= 0 ) STOP 'AEROBULK_MODEL (GPU): the library reported an error'

      !! ---- copy-out for non-contiguous actual arguments
      IF( ALLOCATED(c_QL) ) QL    = c_QL
      IF( ALLOCATED(c_QH) ) QH    = c_QH
      IF( ALLOCATED(c_Tx) ) Tau_x = c_Tx
      IF( ALLOCATED(c_Ty) ) Tau_y = c_Ty
      IF( ALLOCATED(c_Ev) ) Evap  = c_Ev
      IF( ALLOCATED(c_Ts) ) T_s   = c_Ts

      !! ---- mirror the session globals the reference keeps in mod_const
      IF( jt == 1 ) nitend = Nt
      l_use_skin_schemes = ( aerobulk_gpu_get_use_skin() /= 0_c_int )

   END SUBROUTINE AEROBULK_MODEL

END MODULE mod_aerobulk
