/*
 * msed_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, fp64) of MOSSCO's fabm_sediment column solver hot path.
 * It exists to CHECK the CUDA product (mossco_code_b200/csrc) and to be timed as the
 * CPU baseline; nothing in the product path may include, link or call it.
 *
 * Parity status: the in-repo Fortran (solver_library.F90, fabm_sediment_driver.F90,
 * the component's step wrapper) is restated operation by operation and pinned against
 * the only known-answer case the reference holds (src/test/test_Solver.F90) plus
 * closed forms derived from the reference source.  The omexdia_p reaction term lives in
 * an un-vendored, unpinned FABM checkout (external/include/fabm.mk:18) that is absent
 * from /root/reference: for that term this oracle restates SURVEY.md Appendix B and is
 * "PARITY UNPINNED" upstream.
 *
 * All arrays are Fortran order, 0-based here:
 *   3-D  a(i,j,k)   -> a[i + inum*(j + jnum*k)]
 *   4-D  c(i,j,k,n) -> c[i + inum*(j + jnum*(k + knum*n))]
 *   2-D+var b(i,j,n)-> b[i + inum*(j + jnum*n)]
 */
#ifndef MSED_ORACLE_H
#define MSED_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSED_NVAR_OMEXDIA 8

/* ode_solver method ids, solver_library.F90:32-35 */
enum { OSED_EULER = 0, OSED_RK4 = 1, OSED_ADAPTIVE_EULER = 2, OSED_RK4_38 = 3 };

/* reaction model selector (stands in for the FABM model tree, fabm_sediment_driver.F90:328-350) */
enum { OSED_MODEL_OMEXDIA_P = 0, OSED_MODEL_NONE = 1 };

/* hzg_omexdia_p parameters, examples/standalone/omexdia_p/fabm_sed.nml:51-77 (rates in d-1) */
typedef struct {
    double rLabile, rSemilabile, NCrLdet, NCrSdet, PAds, PAdsODU, NH3Ads, CprodMax;
    double rnit, ksO2nitri, rODUox, ksO2oduox, ksO2oxic, ksNO3denit, kinO2denit;
    double kinNO3anox, kinO2anox;
    double init[OSED_NVAR_OMEXDIA];    /* ldetC sdetC detP po4 no3 nh3 oxy odu */
    double minimum[OSED_NVAR_OMEXDIA];
} osed_omexdia_params;

/* type_rhs_driver, solver_library.F90:37-49.  get_rhs is the dynamic dispatch slot. */
typedef struct osed_rhs_driver {
    int inum, jnum, knum, nvar;
    double dt_min;               /* 1.d-9 */
    double relative_change_min;  /* -0.9d0 */
    double *conc;                /* (i,j,k,n) */
    int *mask;                   /* (i,j,k), >0 == masked (land) */
    double last_min_dt;          /* 1.e20 */
    int last_min_dt_grid_cell[4];
    int adaptive_solver_diagnostics;
    void (*get_rhs)(struct osed_rhs_driver *self, double *rhs);
    long n_subcycle_warnings;    /* counts the write(0,*) at solver_library.F90:128 */
    int verbose;
} osed_rhs_driver;

/* sed_nml, fabm_sediment_driver.F90:211-231 (defaults) */
typedef struct {
    double diffusivity, bioturbation, porosity_max, porosity_fac, k_par, pom_flux_max;
    double bioturbation_depth, bioturbation_min;
    double bioturb_k_l, bioturb_L1, bioturb_L2, bioturb_beta, bioturb_b, bioturb_dry_density;
    int bioturbation_profile, distributed_pom_flux;
} osed_sed_nml;

/* type_sed, fabm_sediment_driver.F90:69-113 (base must stay the first member) */
typedef struct {
    osed_rhs_driver base;
    /* fabm_sed_grid :41-56 */
    double dzmin;
    double *zi, *zc, *dz, *dzc;       /* (i,j,knum+1), (i,j,knum), (i,j,knum), (i,j,knum-1) */
    double bioturbation, diffusivity;
    int bioturbation_profile;
    double beta, k_l, b, L1, L2;
    double poc_factor[2];
    double *poc_data[2];
    double k_par;
    double *fluxes;                   /* (i,j,nvar)   */
    double *bdys;                     /* (i,j,nvar+1) */
    int bcup_dissolved_variables, bcup_particulate_variables;
    double porosity_fac, pom_flux_max, missing_value;
    double *porosity, *intf_porosity, *bioturbation_factor, *par, *par_surface;
    double *temp3d, *flux_cap, *biomass, *weighted_toc, *diff, *transport;
    double *diag_denit;               /* hzg_omexdia_p_denit */
    int model;
    osed_omexdia_params p;
    int particulate[OSED_NVAR_OMEXDIA];
    int owns_mask;
} osed_sed;

void osed_sed_nml_defaults(osed_sed_nml *nml);
void osed_omexdia_defaults(osed_omexdia_params *p);

/* fabm_sed_grid%init_grid, fabm_sediment_driver.F90:127-177 */
int osed_init_grid(osed_sed *s, int inum, int jnum, int knum, double dzmin);
/* type_sed%initialize :191-388 ; mask may be NULL (== all wet); mask is (i,j,k) */
int osed_initialize(osed_sed *s, const osed_sed_nml *nml, int model,
                    const osed_omexdia_params *p, const int *mask);
/* update_porosity :393-442 */
void osed_update_porosity(osed_sed *s, int from_surface);
/* init_concentrations :449-481 (s->base.conc must be linked) */
void osed_init_concentrations(osed_sed *s);
/* fabm_sed_check_domain :488-545 ; returns 0 ok, >0 on the reference's "stop" conditions */
int osed_check_domain(osed_sed *s);
/* get_rhs :575-717 */
void osed_get_rhs(osed_rhs_driver *self, double *rhs);
/* diff3d :739-825 */
void osed_diff3d(const osed_sed *s, const double *C, const double *Cup, const double *Cdown,
                 const double *fluxup, const double *fluxdown, int BcUp, int BcDown,
                 const double *D, const double *VF, double *Flux, double *dC,
                 const double *flux_cap);
void osed_finalize(osed_sed *s);

/* ode_solver, solver_library.F90:80-189 */
void osed_ode_solver(osed_rhs_driver *d, double dt, int method);

/* component step wrapper, fabm_sediment_component.F90:1715-1732:
 * ode_solver -> check_NaN (:2377-2421) -> clip to minimum.  returns 0, or 1 if NaN found
 * (in which case the clip is NOT applied, as the reference aborts). */
int osed_component_step(osed_sed *s, double dt, int method);
int osed_check_nan(const osed_sed *s);
/* get_boundary_conditions :1865-2030.  temperature (i,j) may be NULL (skip);
 * csurf[n], wz[n] are (i,j) fields or NULL (variable not in import state). */
void osed_get_boundary_conditions(osed_sed *s, const double *temperature,
                                  const double *const *csurf, const double *const *wz);
/* 1-D spin-up, fabm_sediment_component.F90:557-632; conc1d is (1,1,k,n), filled on exit */
void osed_spinup_column(const osed_sed_nml *nml, const osed_omexdia_params *p, int knum,
                        double dzmin, double dt_min, double relative_change_min,
                        const double *bdys1d, const double *fluxes1d, long nsteps,
                        int method, double *conc1d);

/* reaction term for one cell: SURVEY.md Appendix B (frozen spec) */
void osed_omexdia_p_cell(const osed_omexdia_params *p, const double c[OSED_NVAR_OMEXDIA],
                         double temp_celsius, double rate[OSED_NVAR_OMEXDIA], double *denit);

/* CPU-baseline harness: runs nsteps component steps on a jnum-slab decomposition with
 * nthreads OpenMP threads (one tile per thread, like one MPI rank per tile in the
 * reference); conc/bdys/fluxes/porosity_surface are full-domain arrays (porosity_surface
 * may be NULL). Returns wall seconds of the step loop only, <0 on error. */
double osed_bench_tiled(int inum, int jnum, int knum, double dzmin, const osed_sed_nml *nml,
                        const osed_omexdia_params *p, const int *mask2d, double *conc,
                        const double *bdys, const double *fluxes_in, double dt, int method,
                        int nsteps, double dt_min, double relative_change_min,
                        int bcup_dissolved, int nthreads, long *subcycles);

/* fused-loop CPU variant of the same step: one pass over the state per attempt
 * instead of the reference's whole-array passes; same contract and arguments as osed_bench_tiled.
 * Euler / adaptive Euler, bioturbation_profile != 3, no distributed POM flux; <0 otherwise. */
double osed_bench_fused(int inum, int jnum, int knum, double dzmin, const osed_sed_nml *nml,
                        const osed_omexdia_params *p, const int *mask2d, double *conc,
                        const double *bdys, const double *fluxes_in, double dt, int method,
                        int nsteps, double dt_min, double relative_change_min,
                        int bcup_dissolved, int nthreads, long *subcycles);

/* pelagic <-> soil couplers (src/mediators/pelagic_benthic_coupler.F90, benthic_pelagic_coupler.F90) */
void osed_pelagic_benthic_coupler(size_t n2, const double *const in[10], double *csurf, double *wz);
/* oxy_last_cell != 0: the whole-array assignment of :344-349 as written (every cell = the last cell's value) */
void osed_pelagic_benthic_coupler_ex(size_t n2, const double *const in[10], double *csurf, double *wz,
                                     int oxy_last_cell);
/* pelagic_soil_connector Run (src/mediators/pelagic_soil_connector.F90:176-2122); see msed_oracle.c */
void osed_pelagic_soil_connector(size_t n2, const double *const in[13], const double par[9], int head_compat,
                                 double *csurf, double *wz);
void osed_benthic_pelagic_coupler(size_t n2, const double *up, double dinflux_const, double dipflux_const,
                                  double convertN, double NC_fdet, double NC_sdet, double *out);
/* soil_pelagic_connector Run (src/mediators/soil_pelagic_connector.F90:179-981); out(n2,9) */
void osed_soil_pelagic_connector(size_t n2, const double *up, double dinflux_const, double dipflux_const,
                                 double convertN, double convertP, int want_oxygen, int want_odu, double *out);

/* ---- flat handle API for the Python test harness (oracle/msed_oracle.py) ------------------- */
enum { OSEDPY_CONC = 0, OSEDPY_BDYS, OSEDPY_FLUXES, OSEDPY_POROSITY, OSEDPY_INTF_POROSITY,
       OSEDPY_BIOTURBATION_FACTOR, OSEDPY_PAR, OSEDPY_PAR_SURFACE, OSEDPY_TEMP3D, OSEDPY_FLUX_CAP,
       OSEDPY_BIOMASS, OSEDPY_WEIGHTED_TOC, OSEDPY_DENIT, OSEDPY_ZI, OSEDPY_ZC, OSEDPY_DZ,
       OSEDPY_DZC, OSEDPY_TRANSPORT };
void *osedpy_create(int inum, int jnum, int knum, double dzmin, const osed_sed_nml *nml, int model,
                    const osed_omexdia_params *p, const int *mask2d);
void osedpy_destroy(void *h);
double *osedpy_ptr(void *h, int which);
osed_sed *osedpy_sed(void *h);
void osedpy_set_solver(void *h, double dt_min, double relative_change_min, int bcup_dissolved,
                       int diagnostics, int verbose);
void osedpy_get_solver_diag(void *h, double *last_min_dt, int cell[4], long *subcycles,
                            double *bioturbation);
/* generic type_rhs_driver with the test_Solver.F90:40 right-hand side, for the solver KAT */
void osedpy_test_solver(int inum, int jnum, int knum, int nvar, double *conc, double dt, int method,
                        long nsteps);

#ifdef __cplusplus
}
#endif
#endif
