!-----------------------------------------------------------------------------------------------------------------
! crm_module_b200.F90 -- ISO_C_BINDING shim: the reference's  subroutine crm(...)  re-implemented as a call into the
! B200-native library (include/crm_b200.h).  Drop-in for components/cam/src/physics/crm/crm_module.F90:45-56: same
! module name, same public routine, same argument list, so that the one call site
! (components/cam/src/physics/cam/crm_physics.F90:1145-1156) compiles unchanged.  Build: replace crm_module.F90 by
! this file in the CRM Filepath and link -lcrm_b200 (see INTEGRATION.md).
!
! NOT compiled in the development image (no Fortran compiler there or on the GPU box: profiles/r02_fortran_probe_gpu_box.txt);
! tests/test_shim_static.py checks it statically against include/crm_b200.h (member order of every bind(C) type, every c_loc
! target declared TARGET).  Kept deliberately trivial: it only takes the
! address of every member array of the derived types (c_loc) and forwards the hidden module inputs of the
! reference routine (phys_grid lookups, get_nstep, crmdims constants, accelerate_crm namelist) explicitly.
! The derived types are the reference's own, unmodified (crm_input_module, crm_state_module, crm_rad_module,
! crm_output_module, crm_ecpp_output_module).
!-----------------------------------------------------------------------------------------------------------------
module crm_module
  use iso_c_binding
  use params,                 only: crm_rknd
  use crm_input_module,       only: crm_input_type
  use crm_state_module,       only: crm_state_type
  use crm_rad_module,         only: crm_rad_type
  use crm_output_module,      only: crm_output_type
  use crm_ecpp_output_module, only: crm_ecpp_output_type
  implicit none
  private
  public :: crm
  public :: crm_b200_resident, crm_b200_sync_host

  ! Device-resident mode (SURVEY 8(f) item 1; include/crm_b200.h crm_b200_resident_*): crm_state stays in HBM keyed by lchnk.  The first
  ! crm() call of a chunk uploads the pbuf copy (seeded on the first step, crm_physics.F90:912-939, or read from a restart file);
  ! later calls move only crm_input, crm_rad (the host radiation reads / writes it) and crm_output.  The pbuf copy of crm_state is then
  ! stale until crm_b200_sync_host(lchnk, ...) is called -- the integration patch calls it before restart writes and before the
  ! CRM_U/CRM_W/CRM_T/... outfld calls (crm_physics.F90:1214-1231) on history steps (INTEGRATION.md).
  logical, save :: crm_b200_resident = .false.
  integer, parameter :: max_chunks = 4096
  integer, save :: seen_chunk(max_chunks) = -1, nseen = 0

  ! ---- mirrors of the C structs of include/crm_b200.h (member order and types must match exactly) ----
  type, bind(C) :: crm_dims_t
    integer(c_int) :: ncrms, ld, plev, nx, ny, nzm, nx_rad, ny_rad, yes3d
    real(c_double) :: dx, dt, dt_gl
    integer(c_int) :: igstep, micro, use_crm_accel, crm_accel_uv
    real(c_double) :: crm_accel_factor
    integer(c_int) :: twiddle_f32
    integer(c_int) :: esmt
  end type
  type, bind(C) :: crm_input_t
    type(c_ptr) :: zmid, zint, tl, ql, qccl, qiil, ps, pmid, pint, pdel, phis, ul, vl
    type(c_ptr) :: ocnfrac, tau00, wndls, bflxls, fluxu00, fluxv00, fluxt00, fluxq00
    type(c_ptr) :: lat_deg, lon_deg, gcol
    type(c_ptr) :: ul_esmt, vl_esmt      ! SP_ESMT builds only (ABI >= 4)
  end type
  type, bind(C) :: crm_state_t
    type(c_ptr) :: u_wind, v_wind, w_wind, temperature, qt, qp, qn
    type(c_ptr) :: nc, qr, nr, qi, ni, qs, ns, qg, ng, qc
  end type
  type, bind(C) :: crm_rad_t
    type(c_ptr) :: qrad, temperature, qv, qc, qi, cld, nc, ni, qs, ns
  end type
  type, bind(C) :: crm_output_t
    type(c_ptr) :: qcl, qci, qpl, qpi, tk, tkh
    type(c_ptr) :: prec_crm
    type(c_ptr) :: cltot, cllow, clmed, clhgh, precc, precl, precsc, precsl, jt_crm, mx_crm
    type(c_ptr) :: prectend, precstend, taux, tauy, z0m, timing_factor
    type(c_ptr) :: cldtop, qc_mean, qi_mean, qs_mean, qg_mean, qr_mean
    type(c_ptr) :: sltend, qltend, qcltend, qiltend, cld, gicewp, gliqwp
    type(c_ptr) :: mctot, mcup, mcdn, mcuup, mcudn, mu_crm, md_crm, du_crm, eu_crm, ed_crm
    type(c_ptr) :: flux_qt, fluxsgs_qt, tkez, tkesgsz, tkz, flux_u, flux_v, flux_qp, precflux
    type(c_ptr) :: qt_ls, qt_trans, qp_trans, qp_fall, qp_src, qp_evp, t_ls
    type(c_ptr) :: ultend, vltend   ! SPMOMTRANS builds only, else c_null_ptr (ABI >= 2)
    type(c_ptr) :: wvar, aut, acc, evpc, evpr, mlt, sub, dep, con                       ! m2005 builds only (ABI >= 3)
    type(c_ptr) :: nc_mean, ni_mean, ns_mean, ng_mean, nr_mean
    type(c_ptr) :: aut_a, acc_a, evpc_a, evpr_a, mlt_a, sub_a, dep_a, con_a
    type(c_ptr) :: u_tend_esmt, v_tend_esmt   ! SP_ESMT builds only (ABI >= 4)
  end type

  interface
    integer(c_int) function crm_b200_init(device, dims, handle) bind(C, name="crm_b200_init")
      import :: c_int, c_ptr, crm_dims_t
      integer(c_int), value :: device
      type(crm_dims_t), intent(in) :: dims
      type(c_ptr), intent(out) :: handle
    end function
    integer(c_int) function crm_b200_run(handle, dims, cin, cstate, crad, cout) bind(C, name="crm_b200_run")
      import :: c_int, c_ptr, crm_dims_t, crm_input_t, crm_state_t, crm_rad_t, crm_output_t
      type(c_ptr), value :: handle
      type(crm_dims_t),   intent(in)    :: dims
      type(crm_input_t),  intent(in)    :: cin
      type(crm_state_t),  intent(inout) :: cstate
      type(crm_rad_t),    intent(inout) :: crad
      type(crm_output_t), intent(inout) :: cout
    end function
    integer(c_int) function crm_b200_resident_put(handle, lchnk, dims, cstate, crad) bind(C, name="crm_b200_resident_put")
      import :: c_int, c_ptr, crm_dims_t, crm_state_t, crm_rad_t
      type(c_ptr), value :: handle
      integer(c_int), value :: lchnk
      type(crm_dims_t),   intent(in) :: dims
      type(crm_state_t),  intent(in) :: cstate
      type(crm_rad_t),    intent(in) :: crad
    end function
    integer(c_int) function crm_b200_resident_run(handle, lchnk, dims, cin, crad, cout) bind(C, name="crm_b200_resident_run")
      import :: c_int, c_ptr, crm_dims_t, crm_input_t, crm_rad_t, crm_output_t
      type(c_ptr), value :: handle
      integer(c_int), value :: lchnk
      type(crm_dims_t),   intent(in)    :: dims
      type(crm_input_t),  intent(in)    :: cin
      type(crm_rad_t),    intent(inout) :: crad
      type(crm_output_t), intent(inout) :: cout
    end function
    integer(c_int) function crm_b200_resident_get(handle, lchnk, dims, cstate, crad) bind(C, name="crm_b200_resident_get")
      import :: c_int, c_ptr, crm_dims_t, crm_state_t, crm_rad_t
      type(c_ptr), value :: handle
      integer(c_int), value :: lchnk
      type(crm_dims_t),   intent(in)    :: dims
      type(crm_state_t),  intent(inout) :: cstate
      type(crm_rad_t),    intent(inout) :: crad
    end function
    function crm_b200_last_error(handle) bind(C, name="crm_b200_last_error") result(msg)
      import :: c_ptr
      type(c_ptr), value :: handle
      type(c_ptr) :: msg
    end function
  end interface

  type(c_ptr), save :: handle = c_null_ptr     ! one workspace arena per MPI rank (crm() is not re-entrant in the
  integer,     save :: handle_ncrms = 0        ! reference either: module-level state, SPCAM forbids OpenMP)

contains

  subroutine crm(lchnk, icol, ncrms, dt_gl, plev, crm_input, crm_state, crm_rad, crm_ecpp_output, crm_output)
    use shr_kind_mod,    only: r8 => shr_kind_r8
    use ppgrid,          only: pcols
    use phys_grid,       only: get_rlat_p, get_rlon_p, get_gcol_p
    use time_manager,    only: get_nstep
    use crmdims,         only: crm_nx, crm_ny, crm_nz, crm_nx_rad, crm_ny_rad, crm_dx, crm_dt
    use accelerate_crm_mod, only: use_crm_accel, crm_accel_factor, crm_accel_uv   ! crm_accel_uv is PRIVATE in the reference
                                  ! (accelerate_crm.F90:38): the integration patch adds it to the public list (INTEGRATION.md)
    use cam_abortutils,  only: endrun
    integer , intent(in   ) :: lchnk
    integer , intent(in   ) :: ncrms
    integer , intent(in   ) :: icol(ncrms)
    integer , intent(in   ) :: plev
    real(r8), intent(in   ) :: dt_gl
    type(crm_input_type), target, intent(in   ) :: crm_input        ! TARGET: c_loc() of its members (the caller's actual argument
                                                                    ! needs no change: the interface is explicit, module procedure)
    type(crm_state_type), target, intent(inout) :: crm_state
    type(crm_rad_type),   target, intent(inout) :: crm_rad
    type(crm_ecpp_output_type),   intent(inout) :: crm_ecpp_output      ! untouched (no -DECPP)
    type(crm_output_type), target, intent(inout) :: crm_output

    type(crm_dims_t)   :: d
    type(crm_input_t)  :: ci
    type(crm_state_t)  :: cs
    type(crm_rad_t)    :: cr
    type(crm_output_t) :: co
    real(c_double), allocatable, target :: lat_deg(:), lon_deg(:)
    integer(c_int), allocatable, target :: gcol(:)
    integer :: i, ld, rc, device
    character(len=512) :: msg

    ld = size(crm_state%u_wind, 1)                 ! = pcols: the derived types are allocated with pcols rows
    allocate(lat_deg(ld), lon_deg(ld), gcol(ld))
    lat_deg = 0; lon_deg = 0; gcol = 0
    do i = 1, ncrms                                ! crm_module.F90:286-289, 679
      lat_deg(i) = get_rlat_p(lchnk, icol(i)) * 57.296_r8
      lon_deg(i) = get_rlon_p(lchnk, icol(i)) * 57.296_r8
      gcol(i)    = get_gcol_p(lchnk, icol(i))
    end do

    d%ncrms = ncrms;  d%ld = ld;  d%plev = plev
    d%nx = crm_nx;  d%ny = crm_ny;  d%nzm = crm_nz
    d%nx_rad = crm_nx_rad;  d%ny_rad = crm_ny_rad
    d%yes3d = merge(1, 0, crm_ny > 1)
    d%dx = crm_dx;  d%dt = crm_dt;  d%dt_gl = dt_gl
    d%igstep = get_nstep()
#ifdef m2005
    d%micro = 1
#else
    d%micro = 0
#endif
    d%use_crm_accel = merge(1, 0, use_crm_accel)
    d%crm_accel_uv  = merge(1, 0, crm_accel_uv)
    d%crm_accel_factor = crm_accel_factor
    d%twiddle_f32 = 1                              ! reproduce CRM/fft.F90's implicit REAL(4) trig tables
#if defined(SP_ESMT)
    d%esmt = 1
#else
    d%esmt = 0
#endif

    if (.not. c_associated(handle) .or. ncrms > handle_ncrms) then
      d%ncrms = pcols                              ! size the arena for the largest chunk
      device = 0                                   ! one GPU per rank: CUDA_VISIBLE_DEVICES is set by the launcher
      rc = crm_b200_init(device, d, handle)
      if (rc /= 0) call endrun('crm_b200_init: '//trim(c_message(c_null_ptr)))
      handle_ncrms = pcols
      d%ncrms = ncrms
    end if

    ci%zmid = c_loc(crm_input%zmid);  ci%zint = c_loc(crm_input%zint);  ci%tl = c_loc(crm_input%tl)
    ci%ql = c_loc(crm_input%ql);  ci%qccl = c_loc(crm_input%qccl);  ci%qiil = c_loc(crm_input%qiil)
    ci%ps = c_loc(crm_input%ps);  ci%pmid = c_loc(crm_input%pmid);  ci%pint = c_loc(crm_input%pint)
    ci%pdel = c_loc(crm_input%pdel);  ci%phis = c_loc(crm_input%phis)
    ci%ul = c_loc(crm_input%ul);  ci%vl = c_loc(crm_input%vl)
    ci%ocnfrac = c_loc(crm_input%ocnfrac);  ci%tau00 = c_loc(crm_input%tau00);  ci%wndls = c_loc(crm_input%wndls)
    ci%bflxls = c_loc(crm_input%bflxls);  ci%fluxu00 = c_loc(crm_input%fluxu00);  ci%fluxv00 = c_loc(crm_input%fluxv00)
    ci%fluxt00 = c_loc(crm_input%fluxt00);  ci%fluxq00 = c_loc(crm_input%fluxq00)
    ci%lat_deg = c_loc(lat_deg);  ci%lon_deg = c_loc(lon_deg);  ci%gcol = c_loc(gcol)
    ci%ul_esmt = c_null_ptr;  ci%vl_esmt = c_null_ptr
    co%u_tend_esmt = c_null_ptr;  co%v_tend_esmt = c_null_ptr
#if defined(SP_ESMT)
    ci%ul_esmt = c_loc(crm_input%ul_esmt);  ci%vl_esmt = c_loc(crm_input%vl_esmt)          ! crm_input_module.F90:39-42
    co%u_tend_esmt = c_loc(crm_output%u_tend_esmt);  co%v_tend_esmt = c_loc(crm_output%v_tend_esmt)   ! crm_output_module.F90:81-84
#endif

    cs%u_wind = c_loc(crm_state%u_wind);  cs%v_wind = c_loc(crm_state%v_wind);  cs%w_wind = c_loc(crm_state%w_wind)
    cs%temperature = c_loc(crm_state%temperature);  cs%qt = c_loc(crm_state%qt)
    cs%nc = c_null_ptr; cs%qr = c_null_ptr; cs%nr = c_null_ptr; cs%qi = c_null_ptr; cs%ni = c_null_ptr
    cs%qs = c_null_ptr; cs%ns = c_null_ptr; cs%qg = c_null_ptr; cs%ng = c_null_ptr; cs%qc = c_null_ptr
    cs%qp = c_null_ptr; cs%qn = c_null_ptr
#ifdef m2005
    cs%nc = c_loc(crm_state%nc);  cs%qr = c_loc(crm_state%qr);  cs%nr = c_loc(crm_state%nr);  cs%qi = c_loc(crm_state%qi)
    cs%ni = c_loc(crm_state%ni);  cs%qs = c_loc(crm_state%qs);  cs%ns = c_loc(crm_state%ns);  cs%qg = c_loc(crm_state%qg)
    cs%ng = c_loc(crm_state%ng);  cs%qc = c_loc(crm_state%qc)
#else
    cs%qp = c_loc(crm_state%qp);  cs%qn = c_loc(crm_state%qn)
#endif

    cr%qrad = c_loc(crm_rad%qrad);  cr%temperature = c_loc(crm_rad%temperature);  cr%qv = c_loc(crm_rad%qv)
    cr%qc = c_loc(crm_rad%qc);  cr%qi = c_loc(crm_rad%qi);  cr%cld = c_loc(crm_rad%cld)
    cr%nc = c_null_ptr; cr%ni = c_null_ptr; cr%qs = c_null_ptr; cr%ns = c_null_ptr
#ifdef m2005
    cr%nc = c_loc(crm_rad%nc);  cr%ni = c_loc(crm_rad%ni);  cr%qs = c_loc(crm_rad%qs);  cr%ns = c_loc(crm_rad%ns)
#endif

    co%qcl = c_loc(crm_output%qcl);  co%qci = c_loc(crm_output%qci);  co%qpl = c_loc(crm_output%qpl)
    co%qpi = c_loc(crm_output%qpi);  co%tk = c_loc(crm_output%tk);  co%tkh = c_loc(crm_output%tkh)
    co%prec_crm = c_loc(crm_output%prec_crm)
    co%cltot = c_loc(crm_output%cltot);  co%cllow = c_loc(crm_output%cllow);  co%clmed = c_loc(crm_output%clmed)
    co%clhgh = c_loc(crm_output%clhgh);  co%precc = c_loc(crm_output%precc);  co%precl = c_loc(crm_output%precl)
    co%precsc = c_loc(crm_output%precsc);  co%precsl = c_loc(crm_output%precsl)
    co%jt_crm = c_loc(crm_output%jt_crm);  co%mx_crm = c_loc(crm_output%mx_crm)
    co%prectend = c_loc(crm_output%prectend);  co%precstend = c_loc(crm_output%precstend)
    co%taux = c_loc(crm_output%taux);  co%tauy = c_loc(crm_output%tauy);  co%z0m = c_loc(crm_output%z0m)
    co%timing_factor = c_loc(crm_output%timing_factor)
    co%cldtop = c_loc(crm_output%cldtop);  co%qc_mean = c_loc(crm_output%qc_mean);  co%qi_mean = c_loc(crm_output%qi_mean)
    co%qs_mean = c_loc(crm_output%qs_mean);  co%qg_mean = c_loc(crm_output%qg_mean);  co%qr_mean = c_loc(crm_output%qr_mean)
    co%sltend = c_loc(crm_output%sltend);  co%qltend = c_loc(crm_output%qltend);  co%qcltend = c_loc(crm_output%qcltend)
    co%qiltend = c_loc(crm_output%qiltend);  co%cld = c_loc(crm_output%cld);  co%gicewp = c_loc(crm_output%gicewp)
    co%gliqwp = c_loc(crm_output%gliqwp);  co%mctot = c_loc(crm_output%mctot);  co%mcup = c_loc(crm_output%mcup)
    co%mcdn = c_loc(crm_output%mcdn);  co%mcuup = c_loc(crm_output%mcuup);  co%mcudn = c_loc(crm_output%mcudn)
    co%mu_crm = c_loc(crm_output%mu_crm);  co%md_crm = c_loc(crm_output%md_crm);  co%du_crm = c_loc(crm_output%du_crm)
    co%eu_crm = c_loc(crm_output%eu_crm);  co%ed_crm = c_loc(crm_output%ed_crm)
    co%flux_qt = c_loc(crm_output%flux_qt);  co%fluxsgs_qt = c_loc(crm_output%fluxsgs_qt);  co%tkez = c_loc(crm_output%tkez)
    co%tkesgsz = c_loc(crm_output%tkesgsz);  co%tkz = c_loc(crm_output%tkz);  co%flux_u = c_loc(crm_output%flux_u)
    co%flux_v = c_loc(crm_output%flux_v);  co%flux_qp = c_loc(crm_output%flux_qp);  co%precflux = c_loc(crm_output%precflux)
    co%qt_ls = c_loc(crm_output%qt_ls);  co%qt_trans = c_loc(crm_output%qt_trans);  co%qp_trans = c_loc(crm_output%qp_trans)
    co%qp_fall = c_loc(crm_output%qp_fall);  co%qp_src = c_loc(crm_output%qp_src);  co%qp_evp = c_loc(crm_output%qp_evp)
    co%t_ls = c_loc(crm_output%t_ls)
#if defined(SPMOMTRANS)
    co%ultend = c_loc(crm_output%ultend);  co%vltend = c_loc(crm_output%vltend)
#else
    co%ultend = c_null_ptr;  co%vltend = c_null_ptr
#endif

    co%wvar = c_null_ptr; co%aut = c_null_ptr; co%acc = c_null_ptr; co%evpc = c_null_ptr; co%evpr = c_null_ptr
    co%mlt = c_null_ptr; co%sub = c_null_ptr; co%dep = c_null_ptr; co%con = c_null_ptr
    co%nc_mean = c_null_ptr; co%ni_mean = c_null_ptr; co%ns_mean = c_null_ptr; co%ng_mean = c_null_ptr; co%nr_mean = c_null_ptr
    co%aut_a = c_null_ptr; co%acc_a = c_null_ptr; co%evpc_a = c_null_ptr; co%evpr_a = c_null_ptr
    co%mlt_a = c_null_ptr; co%sub_a = c_null_ptr; co%dep_a = c_null_ptr; co%con_a = c_null_ptr
#ifdef m2005
    co%wvar = c_loc(crm_output%wvar);  co%aut = c_loc(crm_output%aut);  co%acc = c_loc(crm_output%acc)     ! crm_output_module.F90:22-31
    co%evpc = c_loc(crm_output%evpc);  co%evpr = c_loc(crm_output%evpr);  co%mlt = c_loc(crm_output%mlt)
    co%sub = c_loc(crm_output%sub);  co%dep = c_loc(crm_output%dep);  co%con = c_loc(crm_output%con)
    co%nc_mean = c_loc(crm_output%nc_mean);  co%ni_mean = c_loc(crm_output%ni_mean);  co%ns_mean = c_loc(crm_output%ns_mean)   ! :60-74
    co%ng_mean = c_loc(crm_output%ng_mean);  co%nr_mean = c_loc(crm_output%nr_mean)
    co%aut_a = c_loc(crm_output%aut_a);  co%acc_a = c_loc(crm_output%acc_a);  co%evpc_a = c_loc(crm_output%evpc_a)
    co%evpr_a = c_loc(crm_output%evpr_a);  co%mlt_a = c_loc(crm_output%mlt_a);  co%sub_a = c_loc(crm_output%sub_a)
    co%dep_a = c_loc(crm_output%dep_a);  co%con_a = c_loc(crm_output%con_a)
#endif

    if (crm_b200_resident) then
      if (.not. any(seen_chunk(1:nseen) == lchnk)) then        ! first crm() call of this chunk: upload the pbuf copy once
        if (nseen >= max_chunks) call endrun('crm (B200): more than max_chunks resident chunks')
        rc = crm_b200_resident_put(handle, int(lchnk, c_int), d, cs, cr)
        if (rc /= 0) call endrun('crm_b200_resident_put: '//trim(c_message(handle)))
        nseen = nseen + 1
        seen_chunk(nseen) = lchnk
      end if
      rc = crm_b200_resident_run(handle, int(lchnk, c_int), d, ci, cr, co)
    else
      rc = crm_b200_run(handle, d, ci, cs, cr, co)
    end if
    if (rc /= 0) then                               ! replaces task_abort()/endrun of kurant.F90:66-75, damping.F90:39-42
      msg = c_message(handle)
      call endrun('crm (B200): '//trim(msg))
    end if
    deallocate(lat_deg, lon_deg, gcol)
  end subroutine crm

  ! Resident mode: refresh the host (pbuf) copy of crm_state of chunk lchnk before it is written to a restart file or to history.
  subroutine crm_b200_sync_host(lchnk, ncrms, crm_state)
    use ppgrid,          only: pver
    use crmdims,         only: crm_nx, crm_ny, crm_nz, crm_nx_rad, crm_ny_rad, crm_dx, crm_dt
    use cam_abortutils,  only: endrun
    integer, intent(in) :: lchnk, ncrms
    type(crm_state_type), target, intent(inout) :: crm_state
    type(crm_dims_t)  :: d
    type(crm_state_t) :: cs
    type(crm_rad_t)   :: cr
    integer :: rc
    if (.not. crm_b200_resident .or. .not. c_associated(handle)) return
    if (.not. any(seen_chunk(1:nseen) == lchnk)) return
    d%ncrms = ncrms;  d%ld = size(crm_state%u_wind, 1);  d%plev = pver
    d%nx = crm_nx;  d%ny = crm_ny;  d%nzm = crm_nz;  d%nx_rad = crm_nx_rad;  d%ny_rad = crm_ny_rad
    d%yes3d = merge(1, 0, crm_ny > 1);  d%dx = crm_dx;  d%dt = crm_dt;  d%dt_gl = crm_dt
    d%igstep = 2;  d%use_crm_accel = 0;  d%crm_accel_uv = 0;  d%crm_accel_factor = 0;  d%twiddle_f32 = 1;  d%esmt = 0
#ifdef m2005
    d%micro = 1
#else
    d%micro = 0
#endif
    cs%u_wind = c_loc(crm_state%u_wind);  cs%v_wind = c_loc(crm_state%v_wind);  cs%w_wind = c_loc(crm_state%w_wind)
    cs%temperature = c_loc(crm_state%temperature);  cs%qt = c_loc(crm_state%qt)
    cs%nc = c_null_ptr; cs%qr = c_null_ptr; cs%nr = c_null_ptr; cs%qi = c_null_ptr; cs%ni = c_null_ptr
    cs%qs = c_null_ptr; cs%ns = c_null_ptr; cs%qg = c_null_ptr; cs%ng = c_null_ptr; cs%qc = c_null_ptr
    cs%qp = c_null_ptr; cs%qn = c_null_ptr
#ifdef m2005
    cs%nc = c_loc(crm_state%nc);  cs%qr = c_loc(crm_state%qr);  cs%nr = c_loc(crm_state%nr);  cs%qi = c_loc(crm_state%qi)
    cs%ni = c_loc(crm_state%ni);  cs%qs = c_loc(crm_state%qs);  cs%ns = c_loc(crm_state%ns);  cs%qg = c_loc(crm_state%qg)
    cs%ng = c_loc(crm_state%ng);  cs%qc = c_loc(crm_state%qc)
#else
    cs%qp = c_loc(crm_state%qp);  cs%qn = c_loc(crm_state%qn)
#endif
    cr%qrad = c_null_ptr; cr%temperature = c_null_ptr; cr%qv = c_null_ptr; cr%qc = c_null_ptr; cr%qi = c_null_ptr; cr%cld = c_null_ptr
    cr%nc = c_null_ptr; cr%ni = c_null_ptr; cr%qs = c_null_ptr; cr%ns = c_null_ptr      ! crm_rad is copied out by every resident run
    rc = crm_b200_resident_get(handle, int(lchnk, c_int), d, cs, cr)
    if (rc /= 0) call endrun('crm_b200_resident_get: '//trim(c_message(handle)))
  end subroutine crm_b200_sync_host

  function c_message(h) result(s)
    type(c_ptr), intent(in) :: h
    character(len=512) :: s
    type(c_ptr) :: p
    character(kind=c_char), pointer :: c(:)
    integer :: n
    s = ' '
    p = crm_b200_last_error(h)
    if (.not. c_associated(p)) return
    call c_f_pointer(p, c, [512])
    do n = 1, 512
      if (c(n) == c_null_char) exit
      s(n:n) = c(n)
    end do
  end function c_message

end module crm_module
