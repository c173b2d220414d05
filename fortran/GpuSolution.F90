!> @brief Second integration level: the device-resident formulate + outer iteration
!!
!! SOURCE ONLY (no Fortran compiler in this image; bind(C) names are checked against
!! include/mf6gpu.h by tests/test_abi.py).
!!
!! GpuNumericalSolutionType extends NumericalSolutionType and replaces the body of sln_ca
!! (src/Solution/NumericalSolution.f90:1287-1327: prepareSolve, the solve(kiter) loop,
!! finalizeSolve) by ONE call to mf6gpu_solution_timestep for solutions that hold a single
!! GWF model whose packages are all inside the accelerated set (DIS/DISV, NPF without XT3D,
!! STO, CHD, WEL, RIV, RCH, GHB, DRN).  Anything else keeps the inherited host path with the
!! first-level GpuSolverType / GpuMatrixType (fortran/GpuSolver.F90).  The model arrays are
!! handed over once in sln_ar (they are borrowed for that call only), stress data every
!! stress period, heads / FLOW-JA-FACE / package rates come back after every time step into
!! the arrays gwf_ot_dv / gwf_ot_flow / gwf_bd read.
module GpuSolutionBindingsModule
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: mf6gpu_sln_settings, mf6gpu_gwf_model, mf6gpu_bnd_package, mf6gpu_step_report
  public :: MF6GPU_MAX_BUDGET_TERMS
  public :: mf6gpu_solution_create, mf6gpu_solution_destroy, mf6gpu_solution_set_packages
  public :: mf6gpu_solution_set_hfb, mf6gpu_solution_set_gnc
  public :: mf6gpu_solution_timestep, mf6gpu_solution_get_x, mf6gpu_solution_set_x
  public :: mf6gpu_solution_get_flowja, mf6gpu_solution_get_simvals, mf6gpu_solution_get_storage

  integer(c_int), parameter :: MF6GPU_MAX_BUDGET_TERMS = 16

  !> IMS NONLINEAR block + OPTIONS (include/mf6gpu_types.h: mf6gpu_sln_settings)
  type, bind(C) :: mf6gpu_sln_settings
    real(c_double) :: dvclose
    integer(c_int32_t) :: mxiter
    integer(c_int32_t) :: nonmeth
    real(c_double) :: theta
    real(c_double) :: akappa
    real(c_double) :: gamma
    real(c_double) :: amomentum
    integer(c_int32_t) :: iallowptc
    integer(c_int32_t) :: numtrack
    real(c_double) :: btol
    real(c_double) :: breduc
    real(c_double) :: res_lim
  end type mf6gpu_sln_settings

  !> one GWF model: the arrays ConnectionsType / GwfNpfType / GwfStoType hold
  type, bind(C) :: mf6gpu_gwf_model
    integer(c_int32_t) :: index_base
    integer(c_int32_t) :: nodes
    integer(c_int32_t) :: nja
    integer(c_int32_t) :: njas
    type(c_ptr) :: ia, ja, jas, isym, ihc
    type(c_ptr) :: cl1, cl2, hwva, top, bot, area
    type(c_ptr) :: ibound, strt
    type(c_ptr) :: k11, k33, icelltype
    integer(c_int32_t) :: icellavg, inewton, inewtonur, iperched, ivarcv, idewatcv, ithickstrt, insto
    type(c_ptr) :: ibotnode
    type(c_ptr) :: ss, sy, iconvert
    integer(c_int32_t) :: istor_coef, iconf_ss, iorig_ss, reserved
    ! NPF anisotropy (hy_eff, gwf-npf.f90:2280-2355): all c_null_ptr = K22 == K11, no rotation
    type(c_ptr) :: k22, angle1, angle2, angle3
    type(c_ptr) :: conn_nx, conn_ny !< unit normal of every connection (lower -> higher cell), [njas]
    ! NPF REWET (sgwf_npf_wetdry / rewet_check, gwf-npf.f90:2061-2223)
    type(c_ptr) :: wetdry
    real(c_double) :: wetfct
    integer(c_int32_t) :: irewet, iwetit, ihdwet, reserved2
  end type mf6gpu_gwf_model

  type, bind(C) :: mf6gpu_bnd_package
    integer(c_int32_t) :: ptype
    integer(c_int32_t) :: nbound
    integer(c_int32_t) :: index_base
    integer(c_int32_t) :: iflowred
    real(c_double) :: flowred
    type(c_ptr) :: nodelist, b1, b2, b3
  end type mf6gpu_bnd_package

  type, bind(C) :: mf6gpu_step_report
    integer(c_int32_t) :: converged, outer_iterations, inner_iterations, nterms
    real(c_double) :: max_dv
    integer(c_int32_t) :: max_dv_loc, npivot_fixes, nbacktracks, reserved
    real(c_double) :: totrin, totrot, pdiffr
    real(c_double) :: term_in(MF6GPU_MAX_BUDGET_TERMS)
    real(c_double) :: term_out(MF6GPU_MAX_BUDGET_TERMS)
    integer(c_int32_t) :: term_id(MF6GPU_MAX_BUDGET_TERMS)
    real(c_double) :: t_formulate, t_linsolve
  end type mf6gpu_step_report

  interface
    function mf6gpu_solution_create(model, sln, ims, handle) &
      bind(C, name="mf6gpu_solution_create") result(rc)
      import :: c_int, c_ptr, mf6gpu_gwf_model, mf6gpu_sln_settings
      type(mf6gpu_gwf_model), intent(in) :: model
      type(mf6gpu_sln_settings), intent(in) :: sln
      type(c_ptr), value :: ims !< c_loc of a mf6gpu_ims_settings
      type(c_ptr), intent(out) :: handle
      integer(c_int) :: rc
    end function
    function mf6gpu_solution_destroy(handle) bind(C, name="mf6gpu_solution_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int) :: rc
    end function
    function mf6gpu_solution_set_packages(handle, npkg, pkgs) &
      bind(C, name="mf6gpu_solution_set_packages") result(rc)
      import :: c_int, c_int32_t, c_ptr, mf6gpu_bnd_package
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: npkg
      type(mf6gpu_bnd_package), intent(in) :: pkgs(*)
      integer(c_int) :: rc
    end function
    function mf6gpu_solution_set_hfb(handle, nhfb, noden, nodem, hydchr, index_base) &
      bind(C, name="mf6gpu_solution_set_hfb") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: nhfb
      integer(c_int32_t), intent(in) :: noden(*), nodem(*)
      real(c_double), intent(in) :: hydchr(*)
      integer(c_int32_t), value :: index_base
      integer(c_int) :: rc
    end function
    function mf6gpu_solution_set_gnc(handle, ngnc, numj, noden, nodem, nodesj, alphasj, index_base) &
      bind(C, name="mf6gpu_solution_set_gnc") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: ngnc, numj
      integer(c_int32_t), intent(in) :: noden(*), nodem(*), nodesj(*)
      real(c_double), intent(in) :: alphasj(*)
      integer(c_int32_t), value :: index_base
      integer(c_int) :: rc
    end function
    function mf6gpu_solution_timestep(handle, kper, kstp, delt, iss, report) &
      bind(C, name="mf6gpu_solution_timestep") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr, mf6gpu_step_report
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: kper, kstp, iss
      real(c_double), value :: delt
      type(mf6gpu_step_report), intent(out) :: report
      integer(c_int) :: rc
    end function
    function mf6gpu_solution_get_x(handle, x) bind(C, name="mf6gpu_solution_get_x") result(rc)
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: x(*)
      integer(c_int) :: rc
    end function
    function mf6gpu_solution_set_x(handle, x) bind(C, name="mf6gpu_solution_set_x") result(rc)
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: x(*)
      integer(c_int) :: rc
    end function
    function mf6gpu_solution_get_flowja(handle, flowja) &
      bind(C, name="mf6gpu_solution_get_flowja") result(rc)
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: flowja(*)
      integer(c_int) :: rc
    end function
    function mf6gpu_solution_get_simvals(handle, cap, simvals, count) &
      bind(C, name="mf6gpu_solution_get_simvals") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: cap
      real(c_double), intent(out) :: simvals(*)
      integer(c_int32_t), intent(out) :: count
      integer(c_int) :: rc
    end function
    function mf6gpu_solution_get_storage(handle, strgss, strgsy) &
      bind(C, name="mf6gpu_solution_get_storage") result(rc)
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: strgss(*), strgsy(*)
      integer(c_int) :: rc
    end function
  end interface
end module GpuSolutionBindingsModule

module GpuNumericalSolutionModule
  use, intrinsic :: iso_c_binding
  use KindModule, only: I4B, DP
  use NumericalSolutionModule, only: NumericalSolutionType
  use NumericalModelModule, only: NumericalModelType, GetNumericalModelFromList
  use GwfModule, only: GwfModelType
  use BndModule, only: BndType, GetBndFromList
  use ChdModule, only: ChdType
  use WelModule, only: WelType
  use RivModule, only: RivType
  use RchModule, only: RchType
  use GhbModule, only: GhbType
  use DrnModule, only: DrnType
  use SimModule, only: store_error
  use ConstantsModule, only: DZERO
  use TdisModule, only: kper, kstp, delt
  use Mf6GpuBindingsModule, only: mf6gpu_ims_settings, mf6gpu_check
  use GpuSolutionBindingsModule
  implicit none
  private
  public :: GpuNumericalSolutionType

  type, extends(NumericalSolutionType) :: GpuNumericalSolutionType
    type(c_ptr) :: handle = c_null_ptr !< mf6gpu_solution*
    class(GwfModelType), pointer :: gwf => null() !< the one model of this solution
    real(DP), dimension(:), allocatable, target :: simvals_all !< package rates, packages concatenated
    real(DP), dimension(:), allocatable, target :: conn_nx !< x component of every connection's unit normal (lower -> higher cell)
    real(DP), dimension(:), allocatable, target :: conn_ny !< y component
    integer(I4B), dimension(:), allocatable, target :: icelltype_user !< icelltype with THICKSTRT cells marked negative again
    real(DP), dimension(:), allocatable, target :: ddrn !< contiguous copy of the DRN drainage-depth auxiliary column (one DRN package)
  contains
    procedure :: fill_connection_normals => gpu_fill_connection_normals
    procedure :: sln_ar => gpu_sln_ar
    procedure :: sln_rp => gpu_sln_rp
    procedure :: sln_ca => gpu_sln_ca
    procedure :: sln_da => gpu_sln_da
  end type GpuNumericalSolutionType

contains

  !> @brief after the host allocate/read: hand the model arrays to the device once
  subroutine gpu_sln_ar(this)
    class(GpuNumericalSolutionType) :: this
    type(mf6gpu_gwf_model) :: m
    type(mf6gpu_sln_settings) :: s
    type(mf6gpu_ims_settings), target :: l
    class(NumericalModelType), pointer :: mp
    !
    call this%NumericalSolutionType%sln_ar()
    mp => GetNumericalModelFromList(this%modellist, 1)
    select type (mp)
    class is (GwfModelType)
      this%gwf => mp
    end select
    associate (g => this%gwf, con => this%gwf%dis%con)
      m%index_base = 1 ! the Fortran arrays are passed unchanged
      m%nodes = g%dis%nodes
      m%nja = con%nja
      m%njas = con%njas
      m%ia = c_loc(con%ia); m%ja = c_loc(con%ja); m%jas = c_loc(con%jas)
      m%isym = c_loc(con%isym); m%ihc = c_loc(con%ihc)
      m%cl1 = c_loc(con%cl1); m%cl2 = c_loc(con%cl2); m%hwva = c_loc(con%hwva)
      m%top = c_loc(g%dis%top); m%bot = c_loc(g%dis%bot); m%area = c_loc(g%dis%area)
      m%ibound = c_loc(g%ibound); m%strt = c_loc(g%ic%strt)
      m%k11 = c_loc(g%npf%k11); m%k33 = c_loc(g%npf%k33); m%icelltype = c_loc(g%npf%icelltype)
      m%icellavg = g%npf%icellavg; m%inewton = g%inewton; m%inewtonur = g%inewtonur
      m%iperched = g%npf%iperched; m%ivarcv = g%npf%ivarcv; m%idewatcv = g%npf%idewatcv
      m%ithickstrt = g%npf%ithickstrt; m%insto = g%insto
      ! prepcheck has already folded THICKSTRT into icelltype (0) and ithickstartflag (gwf-npf.f90:1851-1875); the
      ! device derives the initial saturation itself from a NEGATIVE icelltype, so hand it the user's marking back
      if (g%npf%ithickstrt /= 0) then
        allocate (this%icelltype_user(g%dis%nodes))
        this%icelltype_user = merge(-1, g%npf%icelltype, g%npf%ithickstartflag /= 0)
        m%icelltype = c_loc(this%icelltype_user)
      end if
      m%ibotnode = c_loc(g%npf%ibotnode)
      if (g%insto > 0) then
        m%ss = c_loc(g%sto%ss); m%sy = c_loc(g%sto%sy); m%iconvert = c_loc(g%sto%iconvert)
        m%istor_coef = g%sto%istor_coef; m%iconf_ss = g%sto%iconf_ss; m%iorig_ss = g%sto%iorig_ss
      else
        m%ss = c_null_ptr; m%sy = c_null_ptr; m%iconvert = c_null_ptr
        m%istor_coef = 0; m%iconf_ss = 0; m%iorig_ss = 0
      end if
      m%reserved = 0
      ! anisotropy: K22 and the rotation angles (already in radians, gwf-npf.f90:1689-1728); the connection normals
      ! come from dis%connection_normal evaluated once per upper-triangle connection into this%conn_nx / conn_ny
      m%k22 = c_null_ptr; m%angle1 = c_null_ptr; m%angle2 = c_null_ptr; m%angle3 = c_null_ptr
      m%conn_nx = c_null_ptr; m%conn_ny = c_null_ptr
      if (g%npf%ik22 > 0) m%k22 = c_loc(g%npf%k22)
      if (g%npf%iangle1 > 0) m%angle1 = c_loc(g%npf%angle1)
      if (g%npf%iangle2 > 0) m%angle2 = c_loc(g%npf%angle2)
      if (g%npf%iangle3 > 0) m%angle3 = c_loc(g%npf%angle3)
      m%wetdry = c_null_ptr; m%irewet = 0; m%iwetit = 1; m%ihdwet = 0; m%wetfct = 1.0_DP; m%reserved2 = 0
      if (g%npf%irewet > 0) then
        m%wetdry = c_loc(g%npf%wetdry); m%irewet = 1
        m%wetfct = g%npf%wetfct; m%iwetit = g%npf%iwetit; m%ihdwet = g%npf%ihdwet
      end if
      if (g%npf%ik22 > 0 .or. g%npf%iangle1 > 0) then
        call this%fill_connection_normals()
        m%conn_nx = c_loc(this%conn_nx); m%conn_ny = c_loc(this%conn_ny)
      end if
    end associate
    ! IMS NONLINEAR block, read by the inherited sln_ar
    s%dvclose = this%dvclose; s%mxiter = this%mxiter; s%nonmeth = this%nonmeth
    s%theta = this%theta; s%akappa = this%akappa; s%gamma = this%gamma; s%amomentum = this%amomentum
    s%iallowptc = this%iallowptc; s%numtrack = this%numtrack
    s%btol = this%btol; s%breduc = this%breduc; s%res_lim = this%res_lim
    ! IMS LINEAR block: ImsLinearSettingsType copied field by field (cf. GpuSolver.F90 gpu_initialize)
    l%dvclose = this%linear_settings%dvclose; l%rclose = this%linear_settings%rclose
    l%icnvgopt = this%linear_settings%icnvgopt; l%iter1 = this%linear_settings%iter1
    l%ilinmeth = this%linear_settings%ilinmeth; l%iscl = this%linear_settings%iscl
    l%iord = this%linear_settings%iord; l%north = this%linear_settings%north
    l%relax = this%linear_settings%relax; l%level = this%linear_settings%level
    l%droptol = this%linear_settings%droptol
    l%gpu_ordering = 2 ! MF6GPU_ORDER_BLOCK_MULTICOLOR
    l%reserved = 0
    call mf6gpu_check(mf6gpu_solution_create(m, s, c_loc(l), this%handle))
    ! single-model ghost node correction (GhostNodeType: nodem1, nodem2, nodesj(numjs, nexg), alphasj(numjs, nexg),
    ! GhostNode.f90:25-50): Fortran's (numjs, nexg) storage IS the row-major [nexg][numjs] layout the C side reads;
    ! a nodesj of 0 = no cell, which index_base 1 maps to "none".  Applied explicitly on the device.
    if (this%gwf%ingnc > 0) then
      call mf6gpu_check(mf6gpu_solution_set_gnc(this%handle, int(this%gwf%gnc%nexg, c_int32_t), &
                                                int(this%gwf%gnc%numjs, c_int32_t), this%gwf%gnc%nodem1, &
                                                this%gwf%gnc%nodem2, this%gwf%gnc%nodesj, this%gwf%gnc%alphasj, &
                                                1_c_int32_t))
    end if
  end subroutine gpu_sln_ar

  !> @brief stress period data of every boundary package (after the packages' bnd_rp)
  subroutine gpu_sln_rp(this)
    class(GpuNumericalSolutionType) :: this
    type(mf6gpu_bnd_package), dimension(:), allocatable :: pk
    class(BndType), pointer :: b
    integer(I4B) :: ip, np
    !
    np = this%gwf%bndlist%Count()
    allocate (pk(np))
    do ip = 1, np
      b => GetBndFromList(this%gwf%bndlist, ip)
      pk(ip)%nbound = b%nbound
      pk(ip)%index_base = 1
      pk(ip)%iflowred = 0
      pk(ip)%flowred = 0.0_DP
      pk(ip)%nodelist = c_loc(b%nodelist)
      pk(ip)%b2 = c_null_ptr
      pk(ip)%b3 = c_null_ptr
      ! the packages keep their stress columns as separate contiguous arrays (gwf-chd.f90:26,
      ! gwf-wel.f90:40, gwf-riv.f90:22-24, gwf-rch.f90:28, gwf-ghb.f90:22-23, gwf-drn.f90:26-27):
      ! they are passed unchanged, in the column order of mf6gpu_types.h
      select type (b)
      type is (ChdType)
        pk(ip)%ptype = 1; pk(ip)%b1 = c_loc(b%head)
      type is (WelType)
        pk(ip)%ptype = 2; pk(ip)%b1 = c_loc(b%q)
        pk(ip)%iflowred = b%iflowred; pk(ip)%flowred = b%flowred
      type is (RivType)
        pk(ip)%ptype = 3; pk(ip)%b1 = c_loc(b%stage); pk(ip)%b2 = c_loc(b%cond); pk(ip)%b3 = c_loc(b%rbot)
      type is (RchType)
        pk(ip)%ptype = 4; pk(ip)%b1 = c_loc(b%recharge)
      type is (GhbType)
        pk(ip)%ptype = 5; pk(ip)%b1 = c_loc(b%bhead); pk(ip)%b2 = c_loc(b%cond)
      type is (DrnType)
        pk(ip)%ptype = 6; pk(ip)%b1 = c_loc(b%elev); pk(ip)%b2 = c_loc(b%cond)
        ! drainage depth (AUXDEPTHNAME column, gwf-drn.f90:501-530) and cubic scaling (NEWTON / DEV_CUBIC_SCALING)
        if (b%iauxddrncol > 0) then
          if (.not. allocated(this%ddrn)) allocate (this%ddrn(b%maxbound))
          this%ddrn(1:b%nbound) = b%auxvar(b%iauxddrncol, 1:b%nbound)
          pk(ip)%b3 = c_loc(this%ddrn)
        end if
        pk(ip)%iflowred = b%icubic_scaling
      class default
        call store_error('package '//trim(b%packName)//' is outside the GPU path', terminate=.true.)
      end select
    end do
    call mf6gpu_check(mf6gpu_solution_set_packages(this%handle, int(np, c_int32_t), pk))
    ! horizontal flow barriers of this stress period (GwfHfbType keeps noden / nodem / hydchr, gwf-hfb.f90:33-36;
    ! hfb_rp has read the list); the device does condsat_reset / condsat_modify, hfb_fc and hfb_cq itself
    if (this%gwf%inhfb > 0) then
      call mf6gpu_check(mf6gpu_solution_set_hfb(this%handle, int(this%gwf%hfb%nhfb, c_int32_t), &
                                                this%gwf%hfb%noden, this%gwf%hfb%nodem, this%gwf%hfb%hydchr, &
                                                1_c_int32_t))
    end if
  end subroutine gpu_sln_rp

  !> @brief one time step on the device instead of prepareSolve / solve(kiter) loop / finalizeSolve
  subroutine gpu_sln_ca(this, isgcnvg, isuppress_output)
    class(GpuNumericalSolutionType) :: this
    integer(I4B), intent(inout) :: isgcnvg
    integer(I4B), intent(in) :: isuppress_output
    type(mf6gpu_step_report) :: rep
    integer(c_int32_t) :: nb
    !
    call mf6gpu_check(mf6gpu_solution_timestep(this%handle, kper, kstp, delt, this%gwf%iss, rep))
    this%icnvg = rep%converged
    if (rep%converged == 0) isgcnvg = 0
    this%itertot_timestep = rep%outer_iterations
    ! what gwf_ot_dv / gwf_ot_flow / gwf_bd read afterwards
    call mf6gpu_check(mf6gpu_solution_get_x(this%handle, this%gwf%x))
    call mf6gpu_check(mf6gpu_solution_get_flowja(this%handle, this%gwf%flowja))
    call mf6gpu_check(mf6gpu_solution_get_simvals(this%handle, int(size(this%simvals_all), c_int32_t), &
                                                  this%simvals_all, nb))
    if (this%gwf%insto > 0) then
      call mf6gpu_check(mf6gpu_solution_get_storage(this%handle, this%gwf%sto%strgss, this%gwf%sto%strgsy))
    end if
    ! rep%term_in / term_out / term_id feed model_bdentry (Budget.f90) in package order
  end subroutine gpu_sln_ca

  !> @brief unit normal of every upper-triangle connection, as hy_eff gets it from dis%connection_normal
  !! (gwf-npf.f90:2318-2320; Dis.f90:1039-1085, Disv.f90:979-1018): evaluated once, the geometry is static
  subroutine gpu_fill_connection_normals(this)
    class(GpuNumericalSolutionType) :: this
    integer(I4B) :: n, m, ipos, jj
    real(DP) :: zc
    associate (con => this%gwf%dis%con)
      allocate (this%conn_nx(con%njas), this%conn_ny(con%njas))
      this%conn_nx = DZERO
      this%conn_ny = DZERO
      do n = 1, this%gwf%dis%nodes
        do ipos = con%ia(n) + 1, con%ia(n + 1) - 1
          m = con%ja(ipos)
          if (m < n) cycle
          jj = con%jas(ipos)
          if (con%ihc(jj) == 0) cycle
          call this%gwf%dis%connection_normal(n, m, con%ihc(jj), this%conn_nx(jj), this%conn_ny(jj), zc, ipos)
        end do
      end do
    end associate
  end subroutine gpu_fill_connection_normals

  subroutine gpu_sln_da(this)
    class(GpuNumericalSolutionType) :: this
    call mf6gpu_check(mf6gpu_solution_destroy(this%handle))
    this%handle = c_null_ptr
    call this%NumericalSolutionType%sln_da()
  end subroutine gpu_sln_da

end module GpuNumericalSolutionModule
