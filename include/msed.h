/*
 * msed.h -- C ABI of libmsed_b200.so: the B200-native fabm_sediment column solver.
 *
 * This is the drop-in boundary for MOSSCO's sediment hot path.  Each entry point names the
 * reference interface it replaces (paths relative to the MOSSCO source tree):
 *   src/utilities/solver_library.F90         type_rhs_driver :37-49, ode_solver :80-189
 *   src/drivers/fabm_sediment_driver.F90     type_sed :69-113 and its type-bound procedures
 *   src/components/fabm_sediment_component.F90  Run inner loop :1700-1769,
 *                                            get_boundary_conditions :1865-2030, export :1773-1822
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types.  Fortran binds it with
 *     ISO_C_BINDING (mossco_code_b200/fortran/msed_b200.F90).
 *   - every host array is caller-owned, fp64, Fortran order, exactly the reference's shapes:
 *       conc(inum,jnum,knum,nvar)  bdys(inum,jnum,nvar+1)  fluxes(inum,jnum,nvar)
 *       3-D fields (inum,jnum,knum), 2-D fields (inum,jnum); mask is int32 (inum,jnum), >0 = land.
 *     The library owns all device memory (same layout: variable-major, layer-major,
 *     cell-contiguous, plane stride padded to 128 B).
 *   - one handle == one type_sed instance (one horizontal tile on one GPU).  Not thread safe.
 *   - return value: 0 success; >0 model conditions (MSED_NAN_DETECTED, MSED_BAD_DOMAIN, ...);
 *     <0 usage/CUDA/NCCL errors.  The library never aborts the process and never falls back
 *     to a CPU path: without a CUDA device every compute entry returns MSED_ERR_CUDA.
 */
#ifndef MSED_H
#define MSED_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSED_ABI_VERSION 1
#define MSED_NVAR 8          /* hzg_omexdia_p state variables */
#define MSED_MAX_LAYERS 64   /* knum limit (per-layer tables live in kernel-parameter constant space) */

/* ode_solver method ids, solver_library.F90:32-35 */
#define MSED_EULER 0
#define MSED_RUNGE_KUTTA_4 1
#define MSED_ADAPTIVE_EULER 2
#define MSED_RUNGE_KUTTA_4_38 3

/* reaction model: stands in for the FABM model tree built at fabm_sediment_driver.F90:328-350 */
#define MSED_MODEL_OMEXDIA_P 0   /* hzg_omexdia_p, 8 state variables */
#define MSED_MODEL_NONE 1        /* transport only (base_get_rhs-like zero reaction) */
#define MSED_MODEL_TEST_SOLVER 2 /* rhs(i,j,k,:)=(i+j+k)*1e-8, src/test/test_Solver.F90:40 (KAT) */

/* return codes */
#define MSED_OK 0
#define MSED_NAN_DETECTED 1      /* check_NaN, fabm_sediment_component.F90:2377-2421 */
#define MSED_BAD_DOMAIN 2        /* fabm_sed_check_domain stops, fabm_sediment_driver.F90:503-530 */
#define MSED_ERR_ARG (-1)
#define MSED_ERR_CUDA (-2)
#define MSED_ERR_NCCL (-3)
#define MSED_ERR_ALLOC (-4)
#define MSED_ERR_STATE (-5)

/* msed_get_field selectors: the export_states catalogue, fabm_sediment_driver.F90:877-924 */
#define MSED_FIELD_POROSITY 0
#define MSED_FIELD_LAYER_HEIGHT 1
#define MSED_FIELD_LAYER_CENTER_DEPTH 2
#define MSED_FIELD_TEMPERATURE 3
#define MSED_FIELD_PAR 4
#define MSED_FIELD_BIOMASS 5
#define MSED_FIELD_BIOTURBATION 6
#define MSED_FIELD_WEIGHTED_TOC 7
#define MSED_FIELD_DENIT 8           /* FABM diagnostic hzg_omexdia_p_denit */
#define MSED_FIELD_INTF_POROSITY 9
#define MSED_FIELD_FLUX_CAP 10

typedef struct msed_handle msed_handle;

/* sed_nml (fabm_sediment_driver.F90:211-231), run_nml solver entries
 * (fabm_sediment_component.F90:59-67,83-88) and the hzg_omexdia_p namelist
 * (examples/standalone/omexdia_p/fabm_sed.nml:51-77) in one plain struct. */
typedef struct msed_config {
    int32_t abi_version;          /* MSED_ABI_VERSION */
    int32_t inum, jnum, knum;     /* local tile, type_rhs_driver :38 */
    int32_t device;               /* CUDA ordinal; -1 = current device */
    int32_t model;                /* MSED_MODEL_* */
    double dzmin;                 /* fabm_sed_grid%dzmin */
    /* sed_nml */
    double diffusivity, bioturbation, porosity_max, porosity_fac, k_par, pom_flux_max;
    double bioturbation_depth, bioturbation_min;
    double bioturb_k_l, bioturb_L1, bioturb_L2, bioturb_beta, bioturb_b, bioturb_dry_density;
    int32_t bioturbation_profile, distributed_pom_flux;
    /* run_nml */
    double dt_min, relative_change_min;
    int32_t bcup_dissolved_variables, adaptive_solver_diagnostics;
    /* hzg_omexdia_p (rates per day, as in the namelist) */
    double rLabile, rSemilabile, NCrLdet, NCrSdet, PAds, PAdsODU, NH3Ads, CprodMax;
    double rnit, ksO2nitri, rODUox, ksO2oduox, ksO2oxic, ksNO3denit, kinO2denit;
    double kinNO3anox, kinO2anox;
    double initial_value[MSED_NVAR]; /* ldetC sdetC detP po4 no3 nh3 oxy odu */
    double minimum[MSED_NVAR];   /* state_variables(n)%minimum, each >= 0 (msed_create returns MSED_ERR_ARG otherwise) */
    /* tile origin in the global grid (only used by MSED_MODEL_TEST_SOLVER and minloc reporting) */
    int32_t i_offset, j_offset;
} msed_config;

/* what a stepping call did; mirrors type_rhs_driver diagnostics (solver_library.F90:44-46) */
typedef struct msed_step_info {
    int64_t steps_done;           /* completed ode_solver calls */
    int64_t rhs_evaluations;      /* get_rhs calls issued (attempts or RK stages) */
    int64_t subcycle_warnings;    /* "solver subcycles" events, solver_library.F90:128 */
    double last_min_dt;           /* :44,:132 */
    int32_t last_min_dt_grid_cell[4]; /* :45,:133 (1-based i,j,k,n; -99 when unset) */
    int32_t nan_detected;
    double kernel_ms;             /* device time of the stepping kernels (CUDA events) */
    int64_t kernel_launches;      /* launches of this library's kernels in the call */
    int64_t fused_pairs;          /* committed fused launches (msed_set_step_fusion): pairs or chains */
    double fused_ms;              /* part of kernel_ms spent in the fused phase */
    int64_t fused_steps;          /* steps those launches advanced (2 per pair, up to 16 per chain) */
} msed_step_info;

/* ---- lifecycle ---------------------------------------------------------------------------- */
/* fills cfg with the reference defaults (fabm_sediment_driver.F90:217-231,
 * fabm_sediment_component.F90:59-67, fabm_sed.nml:51-77) */
int msed_config_defaults(msed_config *cfg);
/* fabm_sed_grid%init_grid (:127-177) + type_sed%initialize (:191-388): grid, porosity,
 * flux_cap, bioturbation_factor, scratch; allocates device state (conc zeroed, mask all wet) */
int msed_create(const msed_config *cfg, msed_handle **out);
/* type_sed%finalize (:723-732) + component Finalize deallocation */
int msed_destroy(msed_handle *h);
const char *msed_last_error(const msed_handle *h);
/* library identification: "msed_b200 <abi> sm_100a" */
const char *msed_version(void);
/* sizeof(msed_config) for what==0, sizeof(msed_step_info) for what==1: lets FFI bindings verify
 * their struct layout at load time */
size_t msed_sizeof(int what);

/* ---- static fields ------------------------------------------------------------------------ */
/* grid%zi(knum+1), zc(knum), dz(knum), dzc(knum-1) of init_grid (horizontally uniform); any may be NULL */
int msed_get_grid(const msed_handle *h, double *zi, double *zc, double *dz, double *dzc);
/* sed%mask from ESMF_GRIDITEM_MASK (component :489-502): mask2d>0 = masked. Applies the
 * side effects of check_domain/update_porosity: porosity=1 (:431,:541) and conc=1e20 (:535) there */
int msed_set_mask(msed_handle *h, const int32_t *mask2d);
/* sed%porosity(:,:,:) direct assignment (ReadRestart, component :1462-1470) */
int msed_set_porosity(msed_handle *h, const double *porosity3d);
/* update_porosity(from_surface=.true.) after porosity(:,:,1)=porosity_at_soil_surface
 * (component :1642-1643; driver :393-442) */
int msed_update_porosity_from_surface(msed_handle *h, const double *porosity_surface2d);
/* sed%par_surface, component :1596 */
int msed_set_par_surface(msed_handle *h, const double *par_surface2d);
/* fabm_sed_check_domain, driver :488-545 */
int msed_check_domain(msed_handle *h);

/* ---- state -------------------------------------------------------------------------------- */
/* init_concentrations, driver :449-481 */
int msed_init_concentrations(msed_handle *h);
int msed_set_state(msed_handle *h, const double *conc);   /* sed%conc => conc */
int msed_get_state(msed_handle *h, double *conc);
/* apply a (1,1,knum,nvar) profile to every unmasked column, component :628-632 */
int msed_set_state_from_column(msed_handle *h, const double *conc1d);

/* ---- boundary ----------------------------------------------------------------------------- */
/* sed%bdys => bdys ; sed%fluxes => fluxes (component :1666-1667, main.F90:121-122) */
int msed_set_boundary(msed_handle *h, const double *bdys, const double *fluxes);
/* get_boundary_conditions on device, component :1865-2030.  temperature, csurf[n], wz[n] are
 * (inum,jnum) import fields; NULL entries mean "field not in the import state" */
int msed_get_boundary_conditions(msed_handle *h, const double *temperature2d,
                                 const double *const *csurf, const double *const *wz);
int msed_get_boundary(msed_handle *h, double *bdys, double *fluxes);
/* sed%fluxes after stepping (dissolved entries = intFlux(:,:,1), driver :692) */
int msed_get_fluxes(msed_handle *h, double *fluxes);
/* <var>_upward_flux_at_soil_surface = -fluxes(:,:,n), component :1819 */
int msed_get_upward_fluxes(msed_handle *h, double *upward_fluxes);
/* export_states(n)%data / diagnostics for <name>_in_soil, component :1773-1822 */
int msed_get_field(msed_handle *h, int which, double *out3d);

/* The <var>_in_soil write-back of the component (fabm_sediment_component.F90:1773-1822 copies every export
 * state to its ESMF field every Run; consumers -- output, restart -- read it at their own cadence) as an
 * asynchronous export: _begin snapshots the state on the device (a device-to-device copy in stream order) and
 * starts the PCIe copy of the snapshot into conc_host(inum,jnum,knum,nvar) -- pinned memory for a truly
 * asynchronous copy -- on the library's copy stream, so that stepping calls issued afterwards overlap it;
 * _wait blocks until conc_host is complete.  The snapshot buffer (state-sized) is allocated at the first call;
 * if the device has no room for it the copy reads the state directly and the next stepping call waits. */
int msed_export_state_begin(msed_handle *h, double *conc_host);
int msed_export_state_wait(msed_handle *h);

/* ---- the hot path ------------------------------------------------------------------------- */
/* type_sed%get_rhs, driver :575-717 (one RHS evaluation incl. its side effect on sed%fluxes) */
int msed_get_rhs(msed_handle *h, double *rhs);
/* ode_solver(sed, dt, method), solver_library.F90:80-189: exactly one call, no NaN check/clip */
int msed_ode_solver(msed_handle *h, double dt, int method, msed_step_info *info);
/* nsteps iterations of the Run loop body (component :1715-1732): ode_solver -> check_NaN ->
 * clip to state_variables(n)%minimum, fused on device.  Stops at the first NaN step. */
int msed_step(msed_handle *h, double dt, int method, int64_t nsteps, msed_step_info *info);
/* the whole `do while (.not.stopTime)` loop (component :1700-1769): steps of dt until
 * run_seconds are covered, the last one shortened (:1705-1708) */
int msed_run(msed_handle *h, double dt, int method, double run_seconds, msed_step_info *info);
/* One whole Run of the component with HOST import/export buffers, component :1493-1829:
 * get_boundary_conditions(import fields) -> the step loop of msed_run -> upward_fluxes(inum,jnum,nvar)
 * = -fluxes (:1819).  Same results as msed_get_boundary_conditions + msed_run +
 * msed_get_upward_fluxes, but the tile is processed in column chunks for the first and the last
 * attempt so the PCIe transfers overlap the kernels (falls back to the plain sequence for RK methods,
 * small tiles, or when an attempt was rejected). */
int msed_run_exchange(msed_handle *h, double dt, int method, double run_seconds,
                      const double *temperature2d, const double *const *csurf, const double *const *wz,
                      double *upward_fluxes, msed_step_info *info);
/* Generation counters of the import fields of msed_run_exchange: gen[0] temperature, gen[1 + 2n] csurf(n),
 * gen[2 + 2n] wz(n) (MSED_NIMPORT_GEN entries).  A field whose counter and host pointer are the ones of its last
 * upload is not uploaded again by the next msed_run_exchange calls -- its device staging row persists between
 * Runs -- e.g. the constant sinking velocities of a pelagic component (the *_z_velocity_at_soil_surface fields of
 * fabm_sediment_component.F90:1865-2030 are re-read every Run only because ESMF gives no cheaper way to know).
 * The caller bumps a field's counter whenever it changes the field's data.  gen == NULL (the default state):
 * every field is uploaded every Run. */
/* Device-side phase marks of the last pipelined msed_run_exchange, in ms from the moment its first import copy was
 * issued: ms4[0] last import field landed (H2D), ms4[1] last stepping kernel done, ms4[2] from there to the last
 * flux field landed on the host (the exposed D2H tail; negative if the copies were done first), ms4[3] whole
 * span.  A measuring aid for the overlap of transfers and kernels (bench.py e2e_phases). */
int msed_get_exchange_timing(const msed_handle *h, double *ms4);
#define MSED_NIMPORT_GEN (1 + 2 * MSED_NVAR)
int msed_set_import_generations(msed_handle *h, const uint64_t *gen);
/* Step fusion (default on): Euler / adaptive-Euler steps are issued in speculative fused launches that
 * read and write the state once for several accepted sub-steps -- pairs (thread per column, two sub-steps) or,
 * for knum <= 64, chains (warp per column with the state in registers, up to 16 sub-steps; one layer per lane up to
 * 32 layers, two above).  Each call is
 * planned as a definite piece of the reference's attempt sequence (solver_library.F90:104-140), the way the
 * last completed step went: every step accepted at dt on its first attempt, or -- in a sub-cycling episode --
 * every step as the rejected attempts at dt (, dt/4) followed by 4 (16) accepted sub-steps of dt/4 (dt/16); a
 * rejected attempt leaves the state unchanged, so its RHS is the one of the sub-step that follows and costs no
 * pass of its own.  A fused launch is committed only if the reference would have taken exactly the planned
 * decisions (every planned rejection seen, no accepted sub-step that should have been rejected :126, nothing
 * stopped by check_NaN); otherwise the same attempts are redone singly from the untouched state.  Results are
 * bit-identical in every mode.  mode: 0 off, 1 auto (default: chains where they apply, else pairs),
 * 2 pairs only, 3 chains wherever knum allows.  Mode 1 picks chains for tiles of up to 65536 wet columns
 * (environment MSED_CHAIN_MAX_COLS; a fifth of that above 32 layers), where a thread per column cannot fill the GPU. */
int msed_set_step_fusion(msed_handle *h, int mode);
/* Runge-Kutta calls (solver_library.F90:142-185) on tiles too large for chains: stages per launch with a thread per
 * column -- 4 (default: the whole call in one pass over the state, rk_quad_kernel; needs knum >= 5, otherwise 2 is
 * used) or 2 (stages 1+2 and 3+4, seven state passes per call, rk_pair_kernel).  Environment MSED_RK_STAGES.
 * Bit-identical results either way, and to the staged path (step fusion off). */
int msed_set_rk_stages_per_launch(msed_handle *h, int stages);
/* number of column chunks msed_run_exchange uses (0 = choose from the tile size: 1, 4 or 8 -- 8 from 2^19 columns on --, always the
 * asynchronous sequence; 1 = the plain sequence of the three separate calls, no overlap) */
int msed_set_exchange_chunks(msed_handle *h, int nchunks);
/* order in which msed_run_exchange walks a coupling interval that consists of fused pairs only: 0 = step by
 * step (each pair over all chunks, committed before the next), 1 = chunk by chunk (every chunk runs the whole
 * interval as soon as its import fields have landed; one commit at the end, the committed state untouched
 * until then).  Same results either way.  Initial value: environment MSED_EXCHANGE_CHUNK_MAJOR, else 1. */
int msed_set_exchange_order(msed_handle *h, int chunk_major);
/* 1-D pre-simulation (component :557-632) for the handle's configuration; conc1d(1,1,knum,nvar) out.
 * bdys1d(nvar+1), fluxes1d(nvar). Runs a 1x1 tile on the same device. */
int msed_spinup_column(const msed_config *cfg, const double *bdys1d, const double *fluxes1d,
                       int64_t nsteps, int method, double *conc1d, msed_step_info *info);

/* The same pre-simulation for a BATCH of nmembers independent 1-D columns in one launch (warp per member,
 * the column in registers for all nsteps calls, one layer per lane up to 32 layers and two above; no distributed POM flux -- other configurations
 * loop over msed_spinup_column).  Members share cfg's grid and sed_nml and differ in bdys1d(nmembers,nvar+1),
 * fluxes1d(nmembers,nvar) (Fortran order: member fastest) and, if members != NULL, in their reaction
 * parameters and initial values.  Every member is its own domain, as sed1d is in the reference: the accept
 * test (solver_library.F90:121), last_min_dt and last_min_dt_grid_cell (:130-135) are per member.
 * Outputs: conc1d(nmembers,1,knum,nvar); info[nmembers] (steps_done, rhs_evaluations, subcycle_warnings,
 * last_min_dt, last_min_dt_grid_cell; kernel_ms / kernel_launches of the one launch in every entry), may be
 * NULL.  Bit-identical to msed_spinup_column member by member. */
typedef struct msed_spinup_member {
    double rLabile, rSemilabile, NCrLdet, NCrSdet, PAds, PAdsODU, NH3Ads, CprodMax;
    double rnit, ksO2nitri, rODUox, ksO2oduox, ksO2oxic, ksNO3denit, kinO2denit;
    double kinNO3anox, kinO2anox;                 /* as in msed_config (rates per day) */
    double initial_value[MSED_NVAR];
} msed_spinup_member;
int msed_spinup_batch(const msed_config *cfg, int32_t nmembers, const msed_spinup_member *members,
                      const double *bdys1d, const double *fluxes1d, int64_t nsteps, int method,
                      double *conc1d, msed_step_info *info);

/* ---- benthic-pelagic exchange on device (BASELINE config 5) -------------------------------- */
/* One well-mixed pelagic box per column, resident on the device, so that a coupling step needs no
 * host<->device traffic.  conc2d(inum,jnum,nvar) are the pelagic concentrations of the sediment's
 * eight species, wz2d(inum,jnum,nvar) their z-velocities (only the particulate entries are read),
 * layer_height2d(inum,jnum) the box height, temperature2d(inum,jnum) the bottom-water temperature. */
int msed_pelagic_init(msed_handle *h, const double *conc2d, const double *wz2d,
                      const double *layer_height2d, const double *temperature2d);
int msed_pelagic_get(msed_handle *h, double *conc2d);
/* ncouplings coupling intervals, each: get_boundary_conditions from the pelagic boxes
 * (fabm_sediment_component.F90:1930-2020: temperature, dissolved bdys, particulate fluxes=-C*w) ->
 * msed_run(dt, method, coupling_seconds) -> pelagic conc += upward_flux*coupling_seconds/layer_height
 * where layer_height>0 (src/components/fabm_pelagic_component.F90:2100-2105). */
int msed_coupled_run(msed_handle *h, double dt, int method, double coupling_seconds,
                     int64_t ncouplings, msed_step_info *info);

/* ---- pelagic <-> soil couplers fused onto the device --------------------- */
/* pelagic_benthic_coupler Run, src/mediators/pelagic_benthic_coupler.F90:281-492: bottom-layer
 * pelagic fields (inum,jnum; NULL = field absent from the import state) are converted to the
 * sediment's *_at_soil_surface / *_z_velocity_at_soil_surface fields and fed straight into
 * get_boundary_conditions (component :1865-2030) without leaving the device. */
typedef struct msed_pelagic_state {
    const double *temperature;        /* temperature_in_water                         (required) */
    const double *oxygen;             /* dissolved_oxygen_in_water: <0 = reduced subst. (required) */
    const double *detN;               /* Detritus_Nitrogen_detN_in_water               (required) */
    const double *detN_z_velocity;    /* ..._z_velocity_in_water                       (required) */
    const double *detC;               /* Detritus_Carbon_detC_in_water, NULL: C:N = 106/16 */
    const double *detP;               /* Detritus_Phosphorus_detP_in_water, NULL: detN/16 */
    const double *detP_z_velocity;    /* NULL: detN_z_velocity */
    const double *nitrate, *ammonium; /* NULL: 0.5*DIN each */
    const double *DIN;                /* required when nitrate, ammonium or DIP is NULL */
    const double *DIP;                /* NULL: DIN/16 */
} msed_pelagic_state;
int msed_pelagic_benthic_coupler(msed_handle *h, const msed_pelagic_state *state);

/* benthic_pelagic_coupler Run, src/mediators/benthic_pelagic_coupler.F90:188-287: the sediment's
 * upward bed fluxes recombined into the pelagic model's flux fields.  Output pointers are host
 * arrays (inum,jnum); NULL = field not in the export state.  When nitrate/ammonium are NULL the
 * DIN branch (:228-236) is taken. */
typedef struct msed_benthic_pelagic_params {
    double dinflux_const;   /* :39, namelist; constant DIN boundary flux per year */
    double dipflux_const;   /* <0: dinflux_const/16 (:156 of soil_pelagic_connector, :40 here) */
    double convertN;        /* :43 */
    double NC_fdet, NC_sdet;/* :205-206 (0.20, 0.04) */
} msed_benthic_pelagic_params;
typedef struct msed_pelagic_fluxes {
    double *nitrate, *ammonium, *DIN, *DIP, *detN, *detC, *detP, *oxygen;
} msed_pelagic_fluxes;
int msed_benthic_pelagic_coupler(msed_handle *h, const msed_benthic_pelagic_params *par,
                                 const msed_pelagic_fluxes *out);

/* soil_pelagic_connector Run, src/mediators/soil_pelagic_connector.F90:179-981 (the generic successor
 * of benthic_pelagic_coupler; namelist /soil_pelagic_connector/ :140, defaults :36-41).  Output
 * pointers are host arrays (inum,jnum); NULL = field not in the export state.  nitrate is passed
 * through unscaled (:333-359), ammonium = convertN*NH4 (:409-411), DIN = (NH4+NO3+dinflux_const/year)*
 * convertN (:467-472), DIP = convertP*(PO4+dipflux_const/year) (:529-533).  Oxygen and reduced
 * substances: both wanted -> plain copies (:660-679); only odu -> odu-oxygen (:696-698); only oxygen
 * -> oxygen-odu (:718-720).  detC = sum of the detritus*carbon fluxes (:842-874); detN and detP are
 * sums over import fields named detritus*nitrogen / detritus*phosphorous (:764,:911), of which
 * omexdia_p has none: both are 0. */
typedef struct msed_soil_pelagic_params {
    double dinflux_const;   /* :36, per year */
    double dipflux_const;   /* :37; <0: dinflux_const/16 (:156) */
    double convertN;        /* :41 */
    double convertP;        /* :40 */
} msed_soil_pelagic_params;
typedef struct msed_soil_pelagic_fluxes {
    double *nitrate, *ammonium, *DIN, *DIP, *oxygen, *odu, *detN, *detC, *detP;
} msed_soil_pelagic_fluxes;
int msed_soil_pelagic_connector(msed_handle *h, const msed_soil_pelagic_params *par,
                                const msed_soil_pelagic_fluxes *out);

/* pelagic_soil_connector Run, src/mediators/pelagic_soil_connector.F90:176-2122 (the generic successor of
 * pelagic_benthic_coupler; namelist /pelagic_soil_connector/ :146-148, defaults :38-46): bottom-layer pelagic
 * fields (inum,jnum; NULL = field absent from the import state) -> the sediment's *_at_soil_surface /
 * *_z_velocity_at_soil_surface fields, fed straight into get_boundary_conditions (component :1865-2030)
 * without leaving the device.  Temperature (:351) and PAR (:330) are passed through.
 *   detritus: C:N = detC/(1E-5+detN) (106/16 without detC, :1022-1066) splits detN into the labile and
 *     semilabile carbon pools with the end members NC_ldet, NC_sdet (:1082-1095); the sinking velocity handed
 *     to the sediment is sinking_factor * fac_env * velocity, fac_env = depth^2/(depth^2+half_sedimentation_
 *     depth^2) * half_sedimentation_tke/(tke+half_sedimentation_tke) + sinking_factor_min/sinking_factor, times
 *     1/(1+(detC/critical_detritus)^4) (:1150-1232; each factor only if its field / parameter is there);
 *   detritus P: detP, else convertN*detN/16 (:1521-1559);
 *   ammonium / nitrate: the field if imported, else DIN - the other one, else DIN/2, else the other one
 *     (:1816-1845, :1930-1965), times convertN;  phosphate: convertP*DIP, else convertP*convertN*DIN/16 with DIN
 *     taken as nitrate+ammonium or twice the one that exists (:2040-2110);
 *   oxygen / reduced substances: both imported -> copies; only one -> its positive part and the positive part
 *     of its negative (:800-860).
 * MSED_COMPAT_P2S_HEAD (msed_set_compat) reproduces what the HEAD revision of the file computes where that
 * differs from the above and is defined: the labile/semilabile carbon CONCENTRATION fields are overwritten with
 * sinking_factor*fac_env*detN (the "velocity" block fetches fieldList(1) again, :1291-1295, :1351-1355) and the
 * two carbon velocity fields are never written (here: 0, the value the component creates them with); the
 * phosphorus velocity is sinking_factor*fac_env*detN unless a detP velocity is imported (:1595-1630);
 * phosphate is recomputed from DIN even when DIP is imported (:2092-2094).  (HEAD's oxygen-only branch
 * dereferences the unassociated odu pointer, :849-853: there is nothing to reproduce.) */
typedef struct msed_pelagic_soil_state {
    const double *temperature;        /* temperature_in_water (required) */
    const double *par;                /* photosynthetically_active_radiation_in_water, NULL: par_surface untouched */
    const double *oxygen, *odu;       /* dissolved_oxygen_in_water / dissolved_reduced_substances_in_water */
    const double *detN;               /* Detritus_Nitrogen_detN_in_water (required) */
    const double *detN_z_velocity;    /* (required) */
    const double *detC;               /* Detritus_Carbon_detC_in_water */
    const double *detP, *detP_z_velocity;
    const double *nitrate, *ammonium, *DIN, *DIP;   /* at least one of nitrate, ammonium, DIN */
    const double *water_depth;        /* water_depth_at_soil_surface */
    const double *tke;                /* turbulent_kinetic_energy_at_soil_surface */
} msed_pelagic_soil_state;
typedef struct msed_pelagic_soil_params {
    double sinking_factor, sinking_factor_min, NC_ldet, NC_sdet;
    double half_sedimentation_depth, half_sedimentation_tke, critical_detritus, convertN, convertP;
} msed_pelagic_soil_params;
/* the module defaults, :38-46 (sinking_factor_min, half_sedimentation_depth and critical_detritus are
 * default-real literals there: 0.02, 0.1 and 60.0 rounded to binary32) */
int msed_pelagic_soil_params_defaults(msed_pelagic_soil_params *par);
int msed_pelagic_soil_connector(msed_handle *h, const msed_pelagic_soil_state *state,
                                const msed_pelagic_soil_params *par);

/* Reference quirks that are NOT reproduced by default, switchable for bit parity with reference coupled runs
 * (flags are OR-ed):
 *   MSED_COMPAT_P2B_OXYGEN_LAST_CELL  pelagic_benthic_coupler assigns the whole oxy/odu arrays inside its i,j
 *     loop (pelagic_benthic_coupler.F90:344-349), so every column gets max(0, +-O2) of the tile's LAST cell;
 *     default: per column.
 *   MSED_COMPAT_P2S_HEAD              see msed_pelagic_soil_connector. */
#define MSED_COMPAT_P2B_OXYGEN_LAST_CELL 1
#define MSED_COMPAT_P2S_HEAD 2
int msed_set_compat(msed_handle *h, int flags);

/* ---- whole-domain diagnostics (the optional reductions beside the accept flag) ---- */
/* Sums over the wet columns of the tile -- and, with reduce_over_ranks != 0 and a communicator
 * (msed_comm_init), over all tiles by ncclAllReduce(double, SUM, 2*nvar) -- of the bed flux fluxes(:,:,n)
 * [mmol m-2 s-1 x columns] and of the inventory sum_k conc*porosity*dz [mmol m-2 x columns] per variable.
 * Multiply by the cell area for budgets.  The per-tile sums have a fixed summation order. */
int msed_diagnostics(msed_handle *h, double *bed_flux_sum, double *inventory, int reduce_over_ranks);
/* Checksum of the state over the wet columns that does not depend on how the domain is cut into tiles:
 * out[0] = sum bits(conc)*(2g+1) mod 2^64, out[1] = xor bits(conc), g = index of the cell in the GLOBAL
 * conc(inum,jnum_global,knum,nvar) array.  global_ncol = inum*jnum_global, col_offset = inum*j_offset.  Tile
 * values combine by wrapping addition / xor; a sharded run must reproduce the single-tile pair. */
int msed_state_checksum(msed_handle *h, int64_t global_ncol, int64_t col_offset, uint64_t out[2]);
/* fp64 pipe throughput of `device` measured with a DFMA micro-kernel (8 independent chains per thread, every
 * SM filled): *tflops = 2 x FMA/s.  The denominator of the fp64 roofline bench.py reports for fused launches. */
int msed_measure_fp64_peak(int device, double *tflops);

/* ---- execution control -------------------------------------------------------------------- */
/* all work is enqueued on this cudaStream_t (default: a private non-blocking stream) */
int msed_set_stream(msed_handle *h, void *cuda_stream);
int msed_synchronize(msed_handle *h);
/* device pointers for zero-copy plumbing (torch/NCCL): state [nvar][knum][ld], ld in doubles */
int msed_device_state(msed_handle *h, void **dev_ptr, size_t *ld);
/* same entry points with DEVICE-resident inputs (already in the library's padded layout is not
 * required: plain [n][ncol] contiguous device arrays) */
int msed_set_boundary_device(msed_handle *h, const void *bdys_dev, const void *fluxes_dev);
int msed_get_fluxes_device(msed_handle *h, void *fluxes_dev);

/* ---- multi-GPU (one handle per rank, j-slab tiles, no halo) -------------------------------- */
/* NCCL is bound at run time (dlopen libnccl.so.2); only the accept/reject + NaN flags of the
 * adaptive step (solver_library.F90:121) are reduced.  id is a 128-byte ncclUniqueId. */
int msed_nccl_unique_id(char id[128]);
int msed_comm_init(msed_handle *h, const char id[128], int nranks, int rank);
int msed_comm_destroy(msed_handle *h);
/* split phase for hosts that own the collective themselves (e.g. torch.distributed):
 * msed_step runs attempt kernels and calls hook(user, dev_flags, count, stream) to MAX-reduce
 * `count` int32 device flags across ranks before each accept/reject decision (count = 4, or 40 for a fused
 * group that plans rejected attempts).  With adaptive_solver_diagnostics the minloc of :131-135 is reduced
 * over the tiles as well (NCCL communicator only) and reported in global grid indices (i_offset, j_offset). */
typedef int (*msed_allreduce_hook)(void *user, void *dev_flags_i32, int count, void *cuda_stream);
int msed_set_allreduce_hook(msed_handle *h, msed_allreduce_hook hook, void *user);

#ifdef __cplusplus
}
#endif
#endif /* MSED_H */
