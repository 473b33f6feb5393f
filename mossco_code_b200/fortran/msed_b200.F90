!> @file msed_b200.F90
!! @brief ISO_C_BINDING shim: drop-in `type_sed` / `ode_solver` over libmsed_b200.so
!!
!! This module replaces the BODIES of MOSSCO's sediment hot path while keeping the names and
!! calling sequences `fabm_sediment_component.F90` uses:
!!   - `type_sed` with `grid%init_grid`, `initialize`, `update_porosity`, `init_concentrations`,
!!     `check_domain`, `get_rhs`, `finalize`          (src/drivers/fabm_sediment_driver.F90:69-113)
!!   - `ode_solver(sed, dt, method)`                   (src/utilities/solver_library.F90:80-189)
!!   - `msed_component_run(sed, dt, method, seconds)`  (the do-while loop of
!!     src/components/fabm_sediment_component.F90:1700-1769 incl. check_NaN and the minimum clip)
!! All arrays stay caller-owned Fortran arrays (conc(i,j,k,n), bdys(i,j,n+1), fluxes(i,j,n)); the
!! library owns the device copies in the same layout, so there is no transposition.
!!
!! NOTE: this file cannot be compiled in the development image (no Fortran compiler); it is kept
!! in lock-step with include/msed.h, and every entry below is exercised with the identical
!! C-ABI call sequence from Python (tests/test_gpu_parity.py) and documented in INTEGRATION.md.
module msed_b200

  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: msed_soil_pelagic_connector, msed_pelagic_soil_connector, msed_pelagic_soil_params_defaults, &
            msed_spinup_batch, msed_set_import_generations, msed_export_state_begin, msed_export_state_wait, &
            msed_set_compat, msed_diagnostics, msed_state_checksum

  integer, parameter, public :: rk = c_double
  integer, parameter, public :: MSED_NVAR = 8
  integer(c_int), parameter, public :: MSED_EULER = 0, MSED_RUNGE_KUTTA_4 = 1, &
                                       MSED_ADAPTIVE_EULER = 2, MSED_RUNGE_KUTTA_4_38 = 3
  integer(c_int), parameter, public :: MSED_OK = 0, MSED_NAN_DETECTED = 1, MSED_BAD_DOMAIN = 2

  !> mirrors `struct msed_config` (include/msed.h) member by member
  type, bind(c), public :: msed_config
    integer(c_int32_t) :: abi_version, inum, jnum, knum, device, model
    real(c_double)     :: dzmin
    real(c_double)     :: diffusivity, bioturbation, porosity_max, porosity_fac, k_par, pom_flux_max
    real(c_double)     :: bioturbation_depth, bioturbation_min
    real(c_double)     :: bioturb_k_l, bioturb_L1, bioturb_L2, bioturb_beta, bioturb_b, bioturb_dry_density
    integer(c_int32_t) :: bioturbation_profile, distributed_pom_flux
    real(c_double)     :: dt_min, relative_change_min
    integer(c_int32_t) :: bcup_dissolved_variables, adaptive_solver_diagnostics
    real(c_double)     :: rLabile, rSemilabile, NCrLdet, NCrSdet, PAds, PAdsODU, NH3Ads, CprodMax
    real(c_double)     :: rnit, ksO2nitri, rODUox, ksO2oduox, ksO2oxic, ksNO3denit, kinO2denit
    real(c_double)     :: kinNO3anox, kinO2anox
    real(c_double)     :: initial_value(MSED_NVAR), minimum(MSED_NVAR)
    integer(c_int32_t) :: i_offset, j_offset
  end type

  !> mirrors `struct msed_step_info`
  type, bind(c), public :: msed_step_info
    integer(c_int64_t) :: steps_done, rhs_evaluations, subcycle_warnings
    real(c_double)     :: last_min_dt
    integer(c_int32_t) :: last_min_dt_grid_cell(4)
    integer(c_int32_t) :: nan_detected
    real(c_double)     :: kernel_ms
    integer(c_int64_t) :: kernel_launches
    integer(c_int64_t) :: fused_pairs
    real(c_double)     :: fused_ms
    integer(c_int64_t) :: fused_steps
  end type

  !> mirror `struct msed_soil_pelagic_params` / `msed_soil_pelagic_fluxes`: the soil_pelagic_connector
  !! mediator's Run (src/mediators/soil_pelagic_connector.F90:179-981) on the device.  The flux members
  !! are `c_loc` of the export fields' farrayPtr (c_null_ptr = field not in the export state).
  type, bind(c), public :: msed_soil_pelagic_params
    real(c_double) :: dinflux_const, dipflux_const, convertN, convertP
  end type
  type, bind(c), public :: msed_soil_pelagic_fluxes
    type(c_ptr) :: nitrate, ammonium, DIN, DIP, oxygen, odu, detN, detC, detP
  end type

  !> mirrors `struct msed_spinup_member` / `msed_pelagic_soil_state` / `msed_pelagic_soil_params` (include/msed.h)
  type, bind(c), public :: msed_spinup_member
    real(c_double) :: rLabile, rSemilabile, NCrLdet, NCrSdet, PAds, PAdsODU, NH3Ads, CprodMax
    real(c_double) :: rnit, ksO2nitri, rODUox, ksO2oduox, ksO2oxic, ksNO3denit, kinO2denit
    real(c_double) :: kinNO3anox, kinO2anox
    real(c_double) :: initial_value(MSED_NVAR)
  end type
  type, bind(c), public :: msed_pelagic_soil_state
    type(c_ptr) :: temperature, par, oxygen, odu, detN, detN_z_velocity, detC, detP, detP_z_velocity
    type(c_ptr) :: nitrate, ammonium, DIN, DIP, water_depth, tke
  end type
  type, bind(c), public :: msed_pelagic_soil_params
    real(c_double) :: sinking_factor, sinking_factor_min, NC_ldet, NC_sdet
    real(c_double) :: half_sedimentation_depth, half_sedimentation_tke, critical_detritus, convertN, convertP
  end type
  integer(c_int), parameter, public :: MSED_COMPAT_P2B_OXYGEN_LAST_CELL = 1, MSED_COMPAT_P2S_HEAD = 2

  interface
    integer(c_int) function msed_config_defaults(cfg) bind(c, name='msed_config_defaults')
      import; type(msed_config), intent(out) :: cfg
    end function
    integer(c_int) function msed_create(cfg, h) bind(c, name='msed_create')
      import; type(msed_config), intent(in) :: cfg; type(c_ptr), intent(out) :: h
    end function
    integer(c_int) function msed_destroy(h) bind(c, name='msed_destroy')
      import; type(c_ptr), value :: h
    end function
    integer(c_int) function msed_get_grid(h, zi, zc, dz, dzc) bind(c, name='msed_get_grid')
      import; type(c_ptr), value :: h; real(c_double) :: zi(*), zc(*), dz(*), dzc(*)
    end function
    integer(c_int) function msed_set_mask(h, mask2d) bind(c, name='msed_set_mask')
      import; type(c_ptr), value :: h; integer(c_int32_t), intent(in) :: mask2d(*)
    end function
    integer(c_int) function msed_set_porosity(h, por) bind(c, name='msed_set_porosity')
      import; type(c_ptr), value :: h; real(c_double), intent(in) :: por(*)
    end function
    integer(c_int) function msed_update_porosity_from_surface(h, por2d) &
        bind(c, name='msed_update_porosity_from_surface')
      import; type(c_ptr), value :: h; real(c_double), intent(in) :: por2d(*)
    end function
    integer(c_int) function msed_set_par_surface(h, par2d) bind(c, name='msed_set_par_surface')
      import; type(c_ptr), value :: h; real(c_double), intent(in) :: par2d(*)
    end function
    integer(c_int) function msed_check_domain(h) bind(c, name='msed_check_domain')
      import; type(c_ptr), value :: h
    end function
    integer(c_int) function msed_init_concentrations(h) bind(c, name='msed_init_concentrations')
      import; type(c_ptr), value :: h
    end function
    integer(c_int) function msed_set_state(h, conc) bind(c, name='msed_set_state')
      import; type(c_ptr), value :: h; real(c_double), intent(in) :: conc(*)
    end function
    integer(c_int) function msed_get_state(h, conc) bind(c, name='msed_get_state')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: conc(*)
    end function
    integer(c_int) function msed_set_boundary(h, bdys, fluxes) bind(c, name='msed_set_boundary')
      import; type(c_ptr), value :: h; real(c_double), intent(in) :: bdys(*), fluxes(*)
    end function
    integer(c_int) function msed_get_boundary_conditions(h, temperature, csurf, wz) &
        bind(c, name='msed_get_boundary_conditions')
      import; type(c_ptr), value :: h, temperature; type(c_ptr), intent(in) :: csurf(*), wz(*)
    end function
    integer(c_int) function msed_get_fluxes(h, fluxes) bind(c, name='msed_get_fluxes')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: fluxes(*)
    end function
    integer(c_int) function msed_get_upward_fluxes(h, up) bind(c, name='msed_get_upward_fluxes')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: up(*)
    end function
    integer(c_int) function msed_get_field(h, which, out3d) bind(c, name='msed_get_field')
      import; type(c_ptr), value :: h; integer(c_int), value :: which; real(c_double), intent(out) :: out3d(*)
    end function
    integer(c_int) function msed_get_rhs(h, rhs) bind(c, name='msed_get_rhs')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: rhs(*)
    end function
    integer(c_int) function msed_ode_solver(h, dt, method, info) bind(c, name='msed_ode_solver')
      import; type(c_ptr), value :: h; real(c_double), value :: dt; integer(c_int), value :: method
      type(msed_step_info), intent(out) :: info
    end function
    integer(c_int) function msed_step(h, dt, method, nsteps, info) bind(c, name='msed_step')
      import; type(c_ptr), value :: h; real(c_double), value :: dt; integer(c_int), value :: method
      integer(c_int64_t), value :: nsteps; type(msed_step_info), intent(out) :: info
    end function
    integer(c_int) function msed_run(h, dt, method, run_seconds, info) bind(c, name='msed_run')
      import; type(c_ptr), value :: h; real(c_double), value :: dt, run_seconds
      integer(c_int), value :: method; type(msed_step_info), intent(out) :: info
    end function
    integer(c_int) function msed_run_exchange(h, dt, method, run_seconds, temperature, csurf, wz, &
        upward, info) bind(c, name='msed_run_exchange')
      import; type(c_ptr), value :: h, temperature; real(c_double), value :: dt, run_seconds
      integer(c_int), value :: method; type(c_ptr), intent(in) :: csurf(*), wz(*)
      real(c_double), intent(out) :: upward(*); type(msed_step_info), intent(out) :: info
    end function
    integer(c_int) function msed_nccl_unique_id(id) bind(c, name='msed_nccl_unique_id')
      import; character(kind=c_char) :: id(128)
    end function
    integer(c_int) function msed_comm_init(h, id, nranks, rank) bind(c, name='msed_comm_init')
      import; type(c_ptr), value :: h; character(kind=c_char), intent(in) :: id(128)
      integer(c_int), value :: nranks, rank
    end function
    integer(c_int) function msed_comm_destroy(h) bind(c, name='msed_comm_destroy')
      import; type(c_ptr), value :: h
    end function
    !> text of the last error of a handle (NULL handle: of the last failed msed_create); a C string
    type(c_ptr) function msed_last_error(h) bind(c, name='msed_last_error')
      import; type(c_ptr), value :: h
    end function
    type(c_ptr) function msed_version() bind(c, name='msed_version')
      import
    end function
    !> sizeof(msed_config) / sizeof(msed_step_info): checked against c_sizeof in `initialize`
    integer(c_size_t) function msed_sizeof(what) bind(c, name='msed_sizeof')
      import; integer(c_int), value :: what
    end function
    integer(c_int) function msed_set_state_from_column(h, conc1d) bind(c, name='msed_set_state_from_column')
      import; type(c_ptr), value :: h; real(c_double), intent(in) :: conc1d(*)
    end function
    integer(c_int) function msed_get_boundary(h, bdys, fluxes) bind(c, name='msed_get_boundary')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: bdys(*), fluxes(*)
    end function
    !> speculative fused launches (results are identical in every mode): 0 off, 1 auto (chains of up to 16
    !! steps with the column in registers where knum <= 32 and the tile is small, else pairs), 2 pairs, 3 chains
    integer(c_int) function msed_set_step_fusion(h, mode) bind(c, name='msed_set_step_fusion')
      import; type(c_ptr), value :: h; integer(c_int), value :: mode
    end function
    integer(c_int) function msed_set_exchange_chunks(h, nchunks) bind(c, name='msed_set_exchange_chunks')
      import; type(c_ptr), value :: h; integer(c_int), value :: nchunks
    end function
    integer(c_int) function msed_set_exchange_order(h, chunk_major) bind(c, name='msed_set_exchange_order')
      import; type(c_ptr), value :: h; integer(c_int), value :: chunk_major
    end function
    !> Runge-Kutta calls with a thread per column: 4 stages per launch (default) or 2
    integer(c_int) function msed_set_rk_stages_per_launch(h, stages) bind(c, name='msed_set_rk_stages_per_launch')
      import; type(c_ptr), value :: h; integer(c_int), value :: stages
    end function
    !> 1-D pre-simulation, fabm_sediment_component.F90:557-632
    integer(c_int) function msed_spinup_column(cfg, bdys1d, fluxes1d, nsteps, method, conc1d, info) &
        bind(c, name='msed_spinup_column')
      import; type(msed_config), intent(in) :: cfg; real(c_double), intent(in) :: bdys1d(*), fluxes1d(*)
      integer(c_int64_t), value :: nsteps; integer(c_int), value :: method
      real(c_double), intent(out) :: conc1d(*); type(msed_step_info), intent(out) :: info
    end function
    !> pelagic boxes coupled on the device (config 5): bed-flux update as fabm_pelagic_component.F90:2100-2105
    integer(c_int) function msed_pelagic_init(h, conc2d, wz2d, layer_height2d, temperature2d) &
        bind(c, name='msed_pelagic_init')
      import; type(c_ptr), value :: h
      real(c_double), intent(in) :: conc2d(*), wz2d(*), layer_height2d(*), temperature2d(*)
    end function
    integer(c_int) function msed_pelagic_get(h, conc2d) bind(c, name='msed_pelagic_get')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: conc2d(*)
    end function
    integer(c_int) function msed_coupled_run(h, dt, method, coupling_seconds, ncouplings, info) &
        bind(c, name='msed_coupled_run')
      import; type(c_ptr), value :: h; real(c_double), value :: dt, coupling_seconds
      integer(c_int), value :: method; integer(c_int64_t), value :: ncouplings
      type(msed_step_info), intent(out) :: info
    end function
    integer(c_int) function msed_soil_pelagic_connector(h, par, fluxes_out) &
        bind(c, name='msed_soil_pelagic_connector')
      import; type(c_ptr), value :: h
      type(msed_soil_pelagic_params), intent(in) :: par; type(msed_soil_pelagic_fluxes), intent(in) :: fluxes_out
    end function
    !> the same pre-simulation for nmembers sed1d clones in one launch (members: c_null_ptr or an array of
    !> msed_spinup_member; bdys1d(nmembers,nvar+1), fluxes1d(nmembers,nvar), conc1d(nmembers,1,knum,nvar))
    integer(c_int) function msed_spinup_batch(cfg, nmembers, members, bdys1d, fluxes1d, nsteps, method, conc1d, info) &
        bind(c, name='msed_spinup_batch')
      import; type(msed_config), intent(in) :: cfg; integer(c_int32_t), value :: nmembers; type(c_ptr), value :: members
      real(c_double), intent(in) :: bdys1d(*), fluxes1d(*)
      integer(c_int64_t), value :: nsteps; integer(c_int), value :: method
      real(c_double), intent(out) :: conc1d(*); type(c_ptr), value :: info
    end function
    !> import fields the coupler has not changed since their last upload stay on the device: gen(1) temperature,
    !> gen(2n), gen(2n+1) the surface concentration and z-velocity of variable n; c_null_ptr: upload everything
    integer(c_int) function msed_set_import_generations(h, gen) bind(c, name='msed_set_import_generations')
      import; type(c_ptr), value :: h, gen
    end function
    !> <name>_in_soil write-back (fabm_sediment_component.F90:1773-1822) at an output cadence: begin snapshots the
    !> state on the device and copies it to conc_host under the following Runs, wait completes it
    integer(c_int) function msed_export_state_begin(h, conc_host) bind(c, name='msed_export_state_begin')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: conc_host(*)
    end function
    integer(c_int) function msed_export_state_wait(h) bind(c, name='msed_export_state_wait')
      import; type(c_ptr), value :: h
    end function
    !> pelagic_soil_connector Run, src/mediators/pelagic_soil_connector.F90:176-2122
    integer(c_int) function msed_pelagic_soil_params_defaults(par) bind(c, name='msed_pelagic_soil_params_defaults')
      import; type(msed_pelagic_soil_params), intent(out) :: par
    end function
    integer(c_int) function msed_pelagic_soil_connector(h, state, par) bind(c, name='msed_pelagic_soil_connector')
      import; type(c_ptr), value :: h
      type(msed_pelagic_soil_state), intent(in) :: state; type(msed_pelagic_soil_params), intent(in) :: par
    end function
    integer(c_int) function msed_set_compat(h, flags) bind(c, name='msed_set_compat')
      import; type(c_ptr), value :: h; integer(c_int), value :: flags
    end function
    !> bed-flux sums and inventories per variable over the wet columns (all tiles with reduce_over_ranks /= 0)
    integer(c_int) function msed_diagnostics(h, bed_flux_sum, inventory, reduce_over_ranks) &
        bind(c, name='msed_diagnostics')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: bed_flux_sum(*), inventory(*)
      integer(c_int), value :: reduce_over_ranks
    end function
    integer(c_int) function msed_state_checksum(h, global_ncol, col_offset, out2) bind(c, name='msed_state_checksum')
      import; type(c_ptr), value :: h; integer(c_int64_t), value :: global_ncol, col_offset
      integer(c_int64_t), intent(out) :: out2(2)
    end function
    integer(c_int) function msed_set_stream(h, cuda_stream) bind(c, name='msed_set_stream')
      import; type(c_ptr), value :: h, cuda_stream
    end function
    integer(c_int) function msed_synchronize(h) bind(c, name='msed_synchronize')
      import; type(c_ptr), value :: h
    end function
  end interface

  !> sediment grid (part of type_sed), fabm_sediment_driver.F90:41-56
  type, public :: fabm_sed_grid
    real(rk), dimension(:,:,:), pointer :: zi=>null(), zc=>null(), dz=>null(), dzc=>null()
    integer  :: knum, inum=-1, jnum=-1
    real(rk) :: dzmin
  end type

  !> drop-in for type_sed (fabm_sediment_driver.F90:69-113) + type_rhs_driver (solver_library.F90:37-49)
  type, public :: type_sed
    type(fabm_sed_grid) :: grid
    integer  :: inum, jnum, knum, nvar = MSED_NVAR
    real(rk) :: dt_min = 1.d-9, relative_change_min = -0.9d0
    real(rk) :: last_min_dt = 1.e20
    integer  :: last_min_dt_grid_cell(4) = (/-99,-99,-99,-99/)
    logical  :: adaptive_solver_diagnostics = .false.
    integer  :: bcup_dissolved_variables = 2
    real(rk), dimension(:,:,:,:), pointer :: conc   => null()
    real(rk), dimension(:,:,:),   pointer :: fluxes => null()
    real(rk), dimension(:,:,:),   pointer :: bdys   => null()
    logical,  dimension(:,:,:),   pointer :: mask   => null()
    real(rk), dimension(:,:,:),   pointer :: porosity => null()
    real(rk), dimension(:,:),     pointer :: par_surface => null()
    type(msed_config)    :: cfg
    type(msed_step_info) :: info
    type(c_ptr)          :: handle = c_null_ptr
    logical              :: state_on_device = .false.
  contains
    procedure :: initialize
    procedure :: finalize
    procedure :: init_concentrations
    procedure :: update_porosity
    procedure :: get_rhs
    procedure :: check_domain
  end type

  public :: ode_solver, msed_component_run, msed_sync_state_to_host, msed_export_fluxes

contains

  !> sed%grid%init_grid + sed%initialize(unit): reads sed_nml / run_nml exactly like the reference
  !! (fabm_sediment_driver.F90:211-233) and creates the device instance.
  subroutine initialize(sed, unit, device)
    class(type_sed), intent(inout) :: sed
    integer, intent(in)            :: unit
    integer, intent(in), optional  :: device
    integer  :: bioturbation_profile, rc, i, j
    logical  :: distributed_pom_flux
    real(rk) :: diffusivity, bioturbation, porosity_max, porosity_fac
    real(rk) :: bioturb_k_l, bioturb_L1, bioturb_L2, bioturb_beta, bioturb_b, bioturb_dry_density
    real(rk) :: k_par, bioturbation_depth, bioturbation_min, pom_flux_max
    integer(c_int32_t), allocatable :: mask2d(:,:)
    namelist /sed_nml/ diffusivity, bioturbation_profile, bioturbation, &
          porosity_max, porosity_fac, k_par, distributed_pom_flux, pom_flux_max, &
          bioturbation_depth, bioturbation_min, bioturb_k_l, bioturb_L1, &
          bioturb_L2, bioturb_beta, bioturb_b, bioturb_dry_density

    block  ! the bind(c) types above must have the layout the library was built with
      type(msed_step_info) :: probe
      if (msed_sizeof(0_c_int) /= c_sizeof(sed%cfg) .or. msed_sizeof(1_c_int) /= c_sizeof(probe)) &
        stop 'msed_b200: struct layout differs from libmsed_b200.so (include/msed.h changed?)'
    end block
    rc = msed_config_defaults(sed%cfg)
    diffusivity = sed%cfg%diffusivity; bioturbation = sed%cfg%bioturbation
    bioturbation_profile = sed%cfg%bioturbation_profile
    porosity_max = sed%cfg%porosity_max; porosity_fac = sed%cfg%porosity_fac; k_par = sed%cfg%k_par
    distributed_pom_flux = .false.; pom_flux_max = sed%cfg%pom_flux_max
    bioturbation_depth = sed%cfg%bioturbation_depth; bioturbation_min = sed%cfg%bioturbation_min
    bioturb_k_l = sed%cfg%bioturb_k_l; bioturb_L1 = sed%cfg%bioturb_L1; bioturb_L2 = sed%cfg%bioturb_L2
    bioturb_beta = sed%cfg%bioturb_beta; bioturb_b = sed%cfg%bioturb_b
    bioturb_dry_density = sed%cfg%bioturb_dry_density
    read(unit, nml=sed_nml)

    sed%inum = sed%grid%inum; sed%jnum = sed%grid%jnum; sed%knum = sed%grid%knum
    sed%cfg%inum = sed%inum; sed%cfg%jnum = sed%jnum; sed%cfg%knum = sed%knum
    sed%cfg%dzmin = sed%grid%dzmin
    if (present(device)) sed%cfg%device = device
    sed%cfg%diffusivity = diffusivity; sed%cfg%bioturbation = bioturbation
    sed%cfg%bioturbation_profile = bioturbation_profile
    sed%cfg%porosity_max = porosity_max; sed%cfg%porosity_fac = porosity_fac; sed%cfg%k_par = k_par
    sed%cfg%distributed_pom_flux = merge(1, 0, distributed_pom_flux); sed%cfg%pom_flux_max = pom_flux_max
    sed%cfg%bioturbation_depth = bioturbation_depth; sed%cfg%bioturbation_min = bioturbation_min
    sed%cfg%bioturb_k_l = bioturb_k_l; sed%cfg%bioturb_L1 = bioturb_L1; sed%cfg%bioturb_L2 = bioturb_L2
    sed%cfg%bioturb_beta = bioturb_beta; sed%cfg%bioturb_b = bioturb_b
    sed%cfg%bioturb_dry_density = bioturb_dry_density
    sed%cfg%dt_min = sed%dt_min; sed%cfg%relative_change_min = sed%relative_change_min
    sed%cfg%bcup_dissolved_variables = sed%bcup_dissolved_variables
    sed%cfg%adaptive_solver_diagnostics = merge(1, 0, sed%adaptive_solver_diagnostics)
    ! hzg_omexdia_p parameters: read namelist /hzg_omexdia_p/ from fabm_sed.nml into sed%cfg here

    rc = msed_create(sed%cfg, sed%handle)
    if (rc /= MSED_OK) stop 'msed_create failed'
    if (associated(sed%mask)) then      ! sed%mask set by the component from ESMF_GRIDITEM_MASK (:489-502)
      allocate(mask2d(sed%inum, sed%jnum))
      do j = 1, sed%jnum; do i = 1, sed%inum
        mask2d(i,j) = merge(1, 0, sed%mask(i,j,1))
      end do; end do
      rc = msed_set_mask(sed%handle, mask2d)
      deallocate(mask2d)
    end if
  end subroutine initialize

  subroutine finalize(sed)
    class(type_sed) :: sed
    integer :: rc
    rc = msed_destroy(sed%handle)
    sed%handle = c_null_ptr
  end subroutine finalize

  !> init_concentrations (fabm_sediment_driver.F90:449-481); conc is refreshed on the host too
  subroutine init_concentrations(sed)
    class(type_sed) :: sed
    integer :: rc
    rc = msed_init_concentrations(sed%handle)
    if (associated(sed%conc)) rc = msed_get_state(sed%handle, sed%conc)
    sed%state_on_device = .true.
  end subroutine init_concentrations

  !> update_porosity(from_surface=.true.) (fabm_sediment_driver.F90:393-442)
  subroutine update_porosity(sed, from_surface)
    class(type_sed)   :: sed
    logical, optional :: from_surface
    integer :: rc
    if (present(from_surface)) then
      if (from_surface) rc = msed_update_porosity_from_surface(sed%handle, sed%porosity(:,:,1))
    end if
  end subroutine update_porosity

  subroutine check_domain(sed, rc)
    class(type_sed)   :: sed
    integer, optional :: rc
    integer :: rc_
    rc_ = msed_check_domain(sed%handle)
    if (rc_ == MSED_BAD_DOMAIN) stop 'FATAL sediment domain check failed'
    if (present(rc)) rc = rc_
  end subroutine check_domain

  !> make the device copy current before stepping (host arrays are authoritative until then)
  subroutine push_inputs(sed)
    class(type_sed) :: sed
    integer :: rc
    if (.not. sed%state_on_device) then
      rc = msed_set_state(sed%handle, sed%conc)
      sed%state_on_device = .true.
    end if
    rc = msed_set_boundary(sed%handle, sed%bdys, sed%fluxes)
    if (associated(sed%par_surface)) rc = msed_set_par_surface(sed%handle, sed%par_surface)
  end subroutine push_inputs

  !> get_rhs (fabm_sediment_driver.F90:575-717)
  subroutine get_rhs(rhs_driver, rhs)
    class(type_sed), intent(inout) :: rhs_driver
    real(rk), intent(inout), dimension(:,:,:,:), pointer :: rhs
    integer :: rc
    call push_inputs(rhs_driver)
    rc = msed_get_rhs(rhs_driver%handle, rhs)
    rc = msed_get_fluxes(rhs_driver%handle, rhs_driver%fluxes)
  end subroutine get_rhs

  !> ode_solver(rhs_driver, dt, method) (solver_library.F90:80): one call, state returned to the host
  subroutine ode_solver(rhs_driver, dt, method)
    integer, intent(in)            :: method
    real(rk), intent(in)           :: dt
    class(type_sed), intent(inout) :: rhs_driver
    integer :: rc
    call push_inputs(rhs_driver)
    rc = msed_ode_solver(rhs_driver%handle, dt, int(method, c_int), rhs_driver%info)
    rhs_driver%last_min_dt = rhs_driver%info%last_min_dt
    rhs_driver%last_min_dt_grid_cell = rhs_driver%info%last_min_dt_grid_cell
    rc = msed_get_state(rhs_driver%handle, rhs_driver%conc)
    rc = msed_get_fluxes(rhs_driver%handle, rhs_driver%fluxes)
  end subroutine ode_solver

  !> The whole Run loop (component :1700-1769) on device: `seconds` of integration in steps of dt
  !! (last step shortened), check_NaN and clip after every step.  rc = MSED_NAN_DETECTED maps to
  !! ESMF_RC_VAL_OUTOFRANGE in the component.  conc stays on the device; call
  !! msed_sync_state_to_host before reading sed%conc (export/output/restart).
  subroutine msed_component_run(sed, dt, method, seconds, rc)
    class(type_sed), intent(inout) :: sed
    real(rk), intent(in)  :: dt, seconds
    integer, intent(in)   :: method
    integer, intent(out)  :: rc
    call push_inputs(sed)
    rc = msed_run(sed%handle, dt, int(method, c_int), seconds, sed%info)
    sed%last_min_dt = sed%info%last_min_dt
    sed%last_min_dt_grid_cell = sed%info%last_min_dt_grid_cell
  end subroutine msed_component_run

  subroutine msed_sync_state_to_host(sed)
    class(type_sed), intent(inout) :: sed
    integer :: rc
    rc = msed_get_state(sed%handle, sed%conc)
  end subroutine msed_sync_state_to_host

  !> <var>_upward_flux_at_soil_surface = -fluxes(:,:,n) (component :1819)
  subroutine msed_export_fluxes(sed, upward)
    class(type_sed), intent(inout) :: sed
    real(rk), intent(out) :: upward(:,:,:)
    integer :: rc
    rc = msed_get_upward_fluxes(sed%handle, upward)
    rc = msed_get_fluxes(sed%handle, sed%fluxes)
  end subroutine msed_export_fluxes

end module msed_b200
