/*
 * msed_oracle.c -- TEST INFRASTRUCTURE ONLY (see msed_oracle.h).
 *
 * Plain-C restatement of the reference's CPU algorithm for the fabm_sediment column
 * solver.  Every routine follows the reference's loop and operation order, including its
 * un-fused whole-array pass structure (the structure is what the CPU baseline times).
 * Paths are relative to /root/reference.
 *
 * Compile the parity build with -ffp-contract=off so no FMA contraction changes the
 * rounding the Fortran source order implies.
 */
#define _POSIX_C_SOURCE 200809L
#include "msed_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <malloc.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define I3(s, i, j, k) ((size_t)(i) + (size_t)(s)->base.inum * ((size_t)(j) + (size_t)(s)->base.jnum * (size_t)(k)))
#define N3(s) ((size_t)(s)->base.inum * (s)->base.jnum * (s)->base.knum)
#define N2(s) ((size_t)(s)->base.inum * (s)->base.jnum)

static double *dalloc(size_t n, double v)
{
    double *p = (double *)malloc((n ? n : 1) * sizeof(double));
    if (!p) { fprintf(stderr, "msed_oracle: out of memory\n"); abort(); }
    for (size_t i = 0; i < n; ++i) p[i] = v;
    return p;
}

/* ---- defaults ------------------------------------------------------------------------- */

/* src/drivers/fabm_sediment_driver.F90:217-231 */
void osed_sed_nml_defaults(osed_sed_nml *nml)
{
    nml->bioturbation_profile = 1;
    nml->diffusivity = 0.9;
    nml->bioturbation = 0.9;
    nml->bioturbation_depth = 5.0;
    nml->bioturbation_min = 0.2;
    nml->porosity_max = 0.7;
    nml->porosity_fac = 0.9;
    nml->k_par = 2.0e-3;
    nml->pom_flux_max = 2.0e4;
    nml->bioturb_k_l = 0.11;
    nml->bioturb_L1 = 0.2;
    nml->bioturb_L2 = 0.6;
    nml->bioturb_beta = 0.22;
    nml->bioturb_b = 1.334;
    nml->bioturb_dry_density = 1000.;
    nml->distributed_pom_flux = 0;
}

/* examples/standalone/omexdia_p/fabm_sed.nml:51-77; state order from
 * examples/standalone/omexdia_p/plotbulknutrients.py:57-62 */
void osed_omexdia_defaults(osed_omexdia_params *p)
{
    p->rLabile = 0.043;  p->rSemilabile = 0.001;
    p->NCrLdet = 0.22;   p->NCrSdet = 0.005;
    p->PAds = 0.01;      p->PAdsODU = 70.;
    p->NH3Ads = 0.0;     p->CprodMax = 9600.0;
    p->rnit = 200.;      p->ksO2nitri = 20.;
    p->rODUox = 20.;     p->ksO2oduox = 1.;
    p->ksO2oxic = 3.;    p->ksNO3denit = 1.;
    p->kinO2denit = 70.; p->kinNO3anox = 1.;
    p->kinO2anox = 1.;
    const double init[8] = {4.e3, 4.e3, 4.e1, 10., 20., 40., 100., 100.};
    for (int n = 0; n < 8; ++n) { p->init[n] = init[n]; p->minimum[n] = 0.0; }
}

/* ---- grid ----------------------------------------------------------------------------- */

/* fabm_sed_grid%init_grid, src/drivers/fabm_sediment_driver.F90:127-177 */
int osed_init_grid(osed_sed *s, int inum, int jnum, int knum, double dzmin)
{
    memset(s, 0, sizeof(*s));
    if (inum < 0 || jnum < 0) return 1; /* :139-142 */
    s->base.inum = inum; s->base.jnum = jnum; s->base.knum = knum;
    s->dzmin = dzmin;
    /* :147 */
    double self_fac = 0.18 / ((knum + 1) / 2.0 * dzmin) - 1.0;
    size_t n2 = (size_t)inum * jnum;
    s->dz = dalloc(n2 * knum, 0.0);
    s->zc = dalloc(n2 * knum, 0.0);
    s->zi = dalloc(n2 * (knum + 1), 0.0);
    s->dzc = dalloc(n2 * (knum > 1 ? knum - 1 : 0), 0.0);
    /* :159-166 (k is 1-based in the Fortran expression) */
    for (int k = 1; k <= knum; ++k)
        for (int j = 0; j < jnum; ++j)
            for (int i = 0; i < inum; ++i) {
                size_t c = I3(s, i, j, k - 1), cn = I3(s, i, j, k);
                s->dz[c] = (1.0 + (self_fac - 1.0) * (double)(k - 1) / (double)(knum - 1)) * dzmin;
                s->zc[c] = s->zi[c] + 0.5 * s->dz[c];
                s->zi[cn] = s->zi[c] + s->dz[c];
            }
    /* :168 */
    for (int k = 0; k < knum - 1; ++k)
        for (size_t c = 0; c < n2; ++c)
            s->dzc[c + n2 * k] = s->zc[c + n2 * (k + 1)] - s->zc[c + n2 * k];
    return 0;
}

/* update_porosity, src/drivers/fabm_sediment_driver.F90:393-442 */
void osed_update_porosity(osed_sed *s, int from_surface)
{
    const int K = s->base.knum;
    const size_t n2 = N2(s);
    if (from_surface) {
        for (int k = 1; k < K; ++k) /* :409-413 */
            for (size_t c = 0; c < n2; ++c)
                s->porosity[c + n2 * k] = s->porosity[c] *
                    (1.0 - s->porosity_fac * (s->zc[c + n2 * k] - s->zc[c]));
        for (int k = 0; k < K; ++k) { /* :415-427 */
            for (size_t c = 0; c < n2; ++c)
                s->flux_cap[c + n2 * k] = s->pom_flux_max / 86400.0 *
                    (1.0 - s->porosity[c + n2 * k]) * s->dz[c + n2 * k];
            if (k + 1 > 2)
                for (size_t c = 0; c < n2; ++c)
                    if (s->flux_cap[c + n2 * k] > s->flux_cap[c + n2 * (k - 1)])
                        s->flux_cap[c + n2 * k] = s->flux_cap[c + n2 * (k - 1)];
        }
    }
    /* :431 */
    for (size_t c = 0; c < N3(s); ++c)
        if (s->base.mask[c] > 0) s->porosity[c] = 1.0;
    /* :434-435 */
    for (size_t c = 0; c < n2; ++c) s->intf_porosity[c] = s->porosity[c];
    for (int k = 1; k < K; ++k)
        for (size_t c = 0; c < n2; ++c)
            s->intf_porosity[c + n2 * k] = 0.5 * (s->porosity[c + n2 * (k - 1)] + s->porosity[c + n2 * k]);
}

/* type_sed%initialize, src/drivers/fabm_sediment_driver.F90:191-388 (numeric parts) */
int osed_initialize(osed_sed *s, const osed_sed_nml *nml, int model,
                    const osed_omexdia_params *p, const int *mask)
{
    const int K = s->base.knum;
    const size_t n2 = N2(s), n3 = N3(s);
    s->bioturbation = nml->bioturbation;                   /* :235-238 */
    s->bioturbation_profile = nml->bioturbation_profile;
    s->diffusivity = nml->diffusivity;
    s->k_par = nml->k_par;
    s->bcup_particulate_variables = nml->distributed_pom_flux ? 4 : 1; /* :239-243 */
    s->bcup_dissolved_variables = 2;                       /* :79 */
    s->missing_value = 1.e20;                              /* :85 */
    s->base.dt_min = 1.e-9;                                /* solver_library.F90:40 */
    s->base.relative_change_min = -0.9;
    s->base.last_min_dt = (double)1.e20f;                  /* :44  real-kind literal 1.e20 */
    for (int q = 0; q < 4; ++q) s->base.last_min_dt_grid_cell[q] = -99;
    s->base.adaptive_solver_diagnostics = 0;
    s->base.get_rhs = osed_get_rhs;

    s->base.mask = (int *)calloc(n3 ? n3 : 1, sizeof(int)); /* :250-253 */
    s->owns_mask = 1;
    if (mask) memcpy(s->base.mask, mask, n3 * sizeof(int));

    s->porosity = dalloc(n3, 0.0);                         /* :262-277 */
    s->intf_porosity = dalloc(n3, 0.0);
    s->bioturbation_factor = dalloc(n3, 1.0);
    s->biomass = dalloc(n3, 0.0);
    s->weighted_toc = dalloc(n3, 0.0);
    s->par = dalloc(n3, 0.0);
    s->par_surface = dalloc(n2, 0.0);
    s->flux_cap = dalloc(n3, 0.0);
    s->porosity_fac = nml->porosity_fac;
    s->pom_flux_max = nml->pom_flux_max;
    for (int k = 0; k < K; ++k) {                          /* :278-304 */
        for (size_t c = 0; c < n2; ++c) {
            size_t q = c + n2 * k;
            s->porosity[q] = nml->porosity_max * (1.0 - nml->porosity_fac * s->zc[q]);
            s->flux_cap[q] = nml->pom_flux_max / 86400.0 * (1.0 - s->porosity[q]) * s->dz[q];
        }
        if (k + 1 > 2)
            for (size_t c = 0; c < n2; ++c)
                if (s->flux_cap[c + n2 * k] > s->flux_cap[c + n2 * (k - 1)])
                    s->flux_cap[c + n2 * k] = s->flux_cap[c + n2 * (k - 1)];
        for (size_t c = 0; c < n2; ++c) {
            size_t q = c + n2 * k;
            switch (nml->bioturbation_profile) {
            case 1:
                s->bioturbation_factor[q] = fmax(nml->bioturbation_min / nml->bioturbation,
                    fmax(nml->bioturbation_depth - 100.0 * s->zi[q], 0.0) / nml->bioturbation_depth);
                break;
            case 2:
                s->bioturbation_factor[q] = exp(-100.0 * s->zi[q] / nml->bioturbation_depth);
                break;
            default: break;
            }
        }
    }
    osed_update_porosity(s, 0);                            /* :313 */

    s->model = model;                                      /* :328-354 */
    if (p) s->p = *p; else osed_omexdia_defaults(&s->p);
    s->base.nvar = OSED_NVAR_OMEXDIA;
    /* particulate property: examples/standalone/omexdia_p/main.F90:92-101 */
    const int part[8] = {1, 1, 1, 0, 0, 0, 0, 0};
    for (int n = 0; n < 8; ++n) s->particulate[n] = part[n];

    s->diff = dalloc(n3, nml->diffusivity);                /* :356-370 */
    s->transport = dalloc(n3 * s->base.nvar, 0.0);
    s->temp3d = dalloc(n3, -999.0);
    s->diag_denit = dalloc(n3, 0.0);

    s->k_l = nml->bioturb_k_l; s->L1 = nml->bioturb_L1; s->L2 = nml->bioturb_L2; /* :373-386 */
    s->beta = nml->bioturb_beta; s->b = nml->bioturb_b;
    s->poc_factor[0] = 1.0 / 1.2 * 12.01 / nml->bioturb_dry_density / 1000.0;
    s->poc_factor[1] = 1.0 / 6.0 * 12.01 / nml->bioturb_dry_density / 1000.0;
    s->poc_data[0] = s->poc_data[1] = NULL;
    return 0;
}

/* init_concentrations, src/drivers/fabm_sediment_driver.F90:449-481 */
void osed_init_concentrations(osed_sed *s)
{
    const size_t n3 = N3(s);
    for (int n = 0; n < s->base.nvar; ++n)                 /* :455-458 */
        for (size_t c = 0; c < n3; ++c)
            s->base.conc[c + n3 * n] = s->p.init[n] / s->porosity[c];
    for (size_t c = 0; c < n3; ++c)                        /* :459-468 */
        if (s->base.mask[c] > 0)
            for (int n = 0; n < s->base.nvar; ++n) s->base.conc[c + n3 * n] = s->missing_value;
    s->poc_data[0] = s->base.conc;                         /* :471-479 */
    s->poc_data[1] = s->base.conc + n3;
}

/* fabm_sed_check_domain, src/drivers/fabm_sediment_driver.F90:488-545 */
int osed_check_domain(osed_sed *s)
{
    const size_t n2 = N2(s), n3 = N3(s);
    const int K = s->base.knum;
    for (size_t c = 0; c < n3; ++c) {
        if (!(s->base.mask[c] > 0) && s->porosity[c] <= 0) return 1;   /* :503 */
        if (!(s->base.mask[c] > 0) && s->porosity[c] > 1) return 2;    /* :508 */
    }
    for (int k = 0; k < K - 1; ++k)                                    /* :513 */
        for (size_t c = 0; c < n2; ++c)
            if (!(s->base.mask[c + n2 * k] > 0) && !(s->base.mask[c + n2 * (k + 1)] > 0) &&
                s->dzc[c + n2 * k] <= 0) return 3;
    for (size_t c = 0; c < n3; ++c)                                    /* :523 */
        if (!(s->base.mask[c] > 0) && s->dz[c] < s->dzmin) return 4;
    if (s->base.conc)                                                  /* :532-538 */
        for (size_t c = 0; c < n3; ++c)
            if (s->base.mask[c] > 0)
                for (int n = 0; n < s->base.nvar; ++n) s->base.conc[c + n3 * n] = 1.e20;
    for (size_t c = 0; c < n3; ++c)                                    /* :541 */
        if (s->base.mask[c] > 0) s->porosity[c] = 1.0;
    return 0;
}

void osed_finalize(osed_sed *s)
{
    free(s->zi); free(s->zc); free(s->dz); free(s->dzc);
    free(s->porosity); free(s->intf_porosity); free(s->bioturbation_factor);
    free(s->biomass); free(s->weighted_toc); free(s->par); free(s->par_surface);
    free(s->flux_cap); free(s->diff); free(s->transport); free(s->temp3d); free(s->diag_denit);
    if (s->owns_mask) free(s->base.mask);
    memset(s, 0, sizeof(*s));
}

/* ---- reaction term -------------------------------------------------------------------- */

/* hzg_omexdia_p `do`: source NOT in /root/reference (FABM is cloned at build time,
 * external/include/fabm.mk:18).  Restates SURVEY.md Appendix B, the project's frozen spec. */
void osed_omexdia_p_cell(const osed_omexdia_params *p, const double c[8], double temp_celsius,
                         double rate[8], double *denit)
{
    const double ldetC = c[0], sdetC = c[1], detP = c[2], po4 = c[3];
    const double no3 = c[4], nh3 = c[5], oxy = c[6], odu = c[7];
    const double relaxO2 = 0.04, T0 = 288.15, Q10b = 1.5;
    const double rLabile = p->rLabile / 86400.0, rSemilabile = p->rSemilabile / 86400.0;
    const double rnit = p->rnit / 86400.0, rODUox = p->rODUox / 86400.0;

    double temp_kelvin = temp_celsius + 273.15;
    double E_a = 0.1 * log(Q10b) * T0 * (T0 + 10.0);
    double f_T = exp(-E_a * (1.0 / temp_kelvin - 1.0 / T0));

    double Oxicminlim = oxy / (oxy + p->ksO2oxic + relaxO2 * (nh3 + odu));
    double Denitrilim = (1.0 - oxy / (oxy + p->kinO2denit)) * no3 / (no3 + p->ksNO3denit);
    double Anoxiclim = (1.0 - oxy / (oxy + p->kinO2anox)) * (1.0 - no3 / (no3 + p->kinNO3anox));
    double Rescale = 1.0 / (Oxicminlim + Denitrilim + Anoxiclim);

    double CprodL = rLabile * ldetC;
    double CprodS = rSemilabile * sdetC;
    double Cprod = CprodL + CprodS;
    double cmax = p->CprodMax / 86400.0;
    if (Cprod > cmax) Cprod = cmax;
    double Nprod = CprodL * p->NCrLdet + CprodS * p->NCrSdet;

    double radsP = p->PAds * rSemilabile * po4 * fmax(odu, p->PAdsODU);
    double rP = rLabile * (1.0 - Oxicminlim);
    double Pprod = rP * detP;

    double OxicMin = Cprod * Oxicminlim * Rescale;
    double Denitrific = Cprod * Denitrilim * Rescale;
    double AnoxicMin = Cprod * Anoxiclim * Rescale;

    double Nitri = f_T * rnit * nh3 * oxy / (oxy + p->ksO2nitri + relaxO2 * (ldetC + odu));
    double OduOx = f_T * rODUox * odu * oxy / (oxy + p->ksO2oduox + relaxO2 * (nh3 + ldetC));

    rate[0] = -f_T * CprodL;
    rate[1] = -f_T * CprodS;
    rate[2] = f_T * (radsP - Pprod);
    rate[3] = f_T * (Pprod - radsP);
    rate[4] = -0.8 * Denitrific + Nitri;
    rate[5] = (Nprod - Nitri) / (1.0 + p->NH3Ads);
    rate[6] = -OxicMin - 2.0 * Nitri - OduOx;
    rate[7] = AnoxicMin - OduOx;
    if (denit) *denit = 0.8 * Denitrific;
}

/* ---- transport ------------------------------------------------------------------------ */

/* diff3d, src/drivers/fabm_sediment_driver.F90:739-825.
 * Flux is (i,j,knum+1); masked columns keep dC = 0 and leave Flux untouched. */
void osed_diff3d(const osed_sed *s, const double *C, const double *Cup, const double *Cdown,
                 const double *fluxup, const double *fluxdown, int BcUp, int BcDown,
                 const double *D, const double *VF, double *Flux, double *dC,
                 const double *flux_cap)
{
    const int inum = s->base.inum, jnum = s->base.jnum, K = s->base.knum;
    const size_t n2 = N2(s), n3 = N3(s);
    for (size_t c = 0; c < n3; ++c) dC[c] = 0.0;                        /* :769 */
    for (int j = 0; j < jnum; ++j)
        for (int i = 0; i < inum; ++i) {
            const size_t c = (size_t)i + (size_t)inum * j;
            if (s->base.mask[c] > 0) continue;                          /* :775 */
            for (int k = 1; k < K; ++k)                                 /* :776-778 */
                Flux[c + n2 * k] = -D[c + n2 * k] * (C[c + n2 * k] - C[c + n2 * (k - 1)]) /
                                   s->dzc[c + n2 * (k - 1)];
            if (BcUp == 1) {                                            /* :782-803 */
                Flux[c] = fluxup[c];
            } else if (BcUp == 2) {
                Flux[c] = -D[c] * (C[c] - Cup[c]) / s->dz[c];
            } else if (BcUp == 3) {
                Flux[c] = 0.0;
            } else if (BcUp == 4) {
                Flux[c] = fluxup[c];
                int k = 2; /* 1-based */
                double restflux = Flux[c] - flux_cap[c];
                while (restflux > 0 && k <= K) {
                    Flux[c + n2 * (k - 1)] = Flux[c + n2 * (k - 1)] + restflux;
                    restflux = restflux - flux_cap[c + n2 * (k - 1)];
                    k = k + 1;
                }
                if (k > K) Flux[c + n2 * (K - 1)] = Flux[c + n2 * (K - 1)] + restflux;
            }
            if (BcDown == 1) {                                          /* :806-816 */
                Flux[c + n2 * K] = fluxdown[c];
            } else if (BcDown == 2) {
                Flux[c + n2 * K] = -D[c + n2 * (K - 1)] * (Cdown[c] - C[c + n2 * (K - 1)]) /
                                   s->dz[c + n2 * (K - 1)];
            } else if (BcDown == 3) {
                Flux[c + n2 * K] = 0.0;
            }
            for (int k = 0; k < K; ++k)                                 /* :818-820 */
                dC[c + n2 * k] = (Flux[c + n2 * k] - Flux[c + n2 * (k + 1)]) /
                                 (VF[c + n2 * k] * s->dz[c + n2 * k]);
        }
}

/* ---- right-hand side ------------------------------------------------------------------- */

/* type_sed%get_rhs, src/drivers/fabm_sediment_driver.F90:575-717 */
void osed_get_rhs(osed_rhs_driver *self, double *rhs)
{
    osed_sed *s = (osed_sed *)self;
    const int K = s->base.knum, nvar = s->base.nvar;
    const size_t n2 = N2(s), n3 = N3(s);
    const int bcdown = 3;                                               /* :590 */
    double *conc_insitu = dalloc(n3, 0.0), *f_T = dalloc(n3, 0.0);      /* :584-588 */
    double *weighted_toc = dalloc(n3, 0.0), *intFlux = dalloc(n2 * (K + 1), 0.0);
    double *cumdepth = dalloc(n2, 0.0), *avg_wtoc = dalloc(n2, 0.0);
    double *volumeFraction = dalloc(n3, 0.0), *zeros2d = dalloc(n2, 0.0);

    /* environmental properties :593-605 */
    for (int k = 0; k < K; ++k) {
        /* sum(dz(:,:,1:k-1),dim=3) ; 0 for k==1.  Layers outermost, columns unit-stride -- the order
         * gfortran gives the intrinsic; per column the terms are still added for m = 1..k-1 */
        for (size_t c = 0; c < n2; ++c) cumdepth[c] = 0.0;
        for (int m = 0; m < k; ++m)
            for (size_t c = 0; c < n2; ++c) cumdepth[c] += s->dz[c + n2 * m];
        for (size_t c = 0; c < n2; ++c)
            if (!(s->base.mask[c + n2 * k] > 0)) {
                s->temp3d[c + n2 * k] = s->bdys[c];
                s->par[c + n2 * k] = s->par_surface[c] * exp(-cumdepth[c] / s->k_par);
            }
    }

    if (s->bioturbation_profile == 3) {                                 /* :618-645 */
        for (size_t c = 0; c < n3; ++c) f_T[c] = 1.0;
        s->bioturbation = 1.0;
        for (size_t c = 0; c < n3; ++c) weighted_toc[c] = 0.0;
        for (int q = 0; q < 2; ++q)
            for (size_t c = 0; c < n3; ++c) {
                if (s->base.mask[c] > 0) continue; /* porosity==1 there: skip the 1/0 */
                weighted_toc[c] = weighted_toc[c] +
                    s->poc_factor[q] * s->porosity[c] / (1.0 - s->porosity[c]) * s->poc_data[q][c];
            }
        for (size_t c = 0; c < n2; ++c) {
            double sum = 0.0;
            for (int k = 0; k < K; ++k) sum += s->dz[c + n2 * k] * weighted_toc[c + n2 * k];
            avg_wtoc[c] = sum / cumdepth[c];   /* cumdepth left over from k==K: excludes bottom layer */
        }
        for (size_t c = 0; c < n3; ++c) s->weighted_toc[c] = weighted_toc[c];
        for (int k = 0; k < K; ++k)
            for (size_t c = 0; c < n2; ++c) {
                size_t q = c + n2 * k;
                s->biomass[q] = weighted_toc[q] * exp(s->zc[q] * 100.0 * s->k_l) * avg_wtoc[c] /
                                (s->L1 + s->L2 * exp(s->zc[q] * 200.0 * s->k_l));
            }
        for (size_t c = 0; c < n3; ++c) {
            if (s->base.mask[c] > 0) continue;
            s->bioturbation_factor[c] = s->beta * pow(s->biomass[c], s->b) / weighted_toc[c];
        }
    } else {                                                            /* :648 */
        for (size_t c = 0; c < n3; ++c)
            f_T[c] = 1.0 * exp(-4500.0 * (1.0 / (s->temp3d[c] + 273.0) - (1.0 / 288.0)));
    }

    for (int n = 0; n < nvar; ++n) {                                    /* :651-694 */
        for (size_t c = 0; c < n3; ++c)                                 /* :652-653 */
            s->diff[c] = s->bioturbation * f_T[c] / 86400.0 / 10000.0 *
                         (1.0 - s->intf_porosity[c]) * s->bioturbation_factor[c];
        double *tr = s->transport + n3 * n;
        if (s->particulate[n]) {                                        /* :658-678 */
            int bcup = s->bcup_particulate_variables;
            for (size_t c = 0; c < n3; ++c) conc_insitu[c] = s->base.conc[c + n3 * n] * s->porosity[c];
            for (size_t c = 0; c < n3; ++c) volumeFraction[c] = 1.0 - s->porosity[c];
            osed_diff3d(s, conc_insitu, s->bdys + n2 * (n + 1), zeros2d, s->fluxes + n2 * n, zeros2d,
                        bcup, bcdown, s->diff, volumeFraction, intFlux, tr, s->flux_cap);
            for (size_t c = 0; c < n3; ++c)
                tr[c] = tr[c] * (1.0 - s->porosity[c]) / s->porosity[c];
        } else {                                                        /* :679-693 */
            int bcup = s->bcup_dissolved_variables;
            for (size_t c = 0; c < n3; ++c)
                s->diff[c] = s->diff[c] + (s->diffusivity + s->temp3d[c] * 0.035) *
                                          s->intf_porosity[c] / 86400.0 / 10000.0;
            for (size_t c = 0; c < n3; ++c) conc_insitu[c] = s->base.conc[c + n3 * n];
            osed_diff3d(s, conc_insitu, s->bdys + n2 * (n + 1), zeros2d, s->fluxes + n2 * n, zeros2d,
                        bcup, bcdown, s->diff, s->porosity, intFlux, tr, NULL);
            /* :692 -- masked columns hold undefined Flux in the reference; defined as 0 here */
            for (size_t c = 0; c < n2; ++c)
                s->fluxes[c + n2 * n] = (s->base.mask[c] > 0) ? 0.0 : intFlux[c];
        }
    }

    for (size_t c = 0; c < n3 * nvar; ++c) rhs[c] = 0.0;                /* :696 */
    for (int k = 0; k < K; ++k)                                         /* :697-712 */
        for (int j = 0; j < s->base.jnum; ++j)
            for (int i = 0; i < s->base.inum; ++i) {
                size_t q = I3(s, i, j, k);
                if (!(s->base.mask[q] > 0)) {
                    if (s->model == OSED_MODEL_OMEXDIA_P) {             /* fabm_do :700 */
                        double c8[8], r8[8], denit;
                        for (int n = 0; n < 8; ++n) c8[n] = s->base.conc[q + n3 * n];
                        osed_omexdia_p_cell(&s->p, c8, s->temp3d[q], r8, &denit);
                        for (int n = 0; n < 8; ++n) rhs[q + n3 * n] = r8[n];
                        s->diag_denit[q] = denit;
                    }
                } else {
                    for (int n = 0; n < nvar; ++n) {
                        rhs[q + n3 * n] = 0.0;
                        s->transport[q + n3 * n] = 0.0;
                    }
                }
            }
    for (size_t c = 0; c < n3 * nvar; ++c) rhs[c] = rhs[c] + s->transport[c]; /* :715 */

    free(conc_insitu); free(f_T); free(weighted_toc); free(intFlux);
    free(cumdepth); free(avg_wtoc); free(volumeFraction); free(zeros2d);
}

/* ---- ode_solver ------------------------------------------------------------------------ */

/* Fortran minloc over (c1-c)/c in array order, NaNs skipped as gfortran does */
static void minloc_relchange(const osed_rhs_driver *d, const double *c1, const double *c, int out[4])
{
    size_t n = (size_t)d->inum * d->jnum * d->knum * d->nvar, best = 0;
    int found = 0;
    double bv = 0.0;
    for (size_t q = 0; q < n; ++q) {
        double v = (c1[q] - c[q]) / c[q];
        if (v != v) continue;
        if (!found || v < bv) { bv = v; best = q; found = 1; }
    }
    out[0] = (int)(best % d->inum) + 1; best /= d->inum;
    out[1] = (int)(best % d->jnum) + 1; best /= d->jnum;
    out[2] = (int)(best % d->knum) + 1; best /= d->knum;
    out[3] = (int)best + 1;
}

/* ode_solver, src/utilities/solver_library.F90:80-189 */
void osed_ode_solver(osed_rhs_driver *d, double dt, int method)
{
    const size_t n = (size_t)d->inum * d->jnum * d->knum * d->nvar;
    double *rhs0 = dalloc(n, 0.0), *rhs1 = NULL, *rhs2 = NULL, *rhs3 = NULL;   /* :89-90 */
    double *c1 = dalloc(n, 0.0);
    const double third = 1.0 / 3.0;                                            /* :96 */

    switch (method) {
    case OSED_EULER:                                                           /* :99-102 */
        d->get_rhs(d, rhs0);
        for (size_t q = 0; q < n; ++q) d->conc[q] = d->conc[q] + dt * rhs0[q];
        break;

    case OSED_ADAPTIVE_EULER: {                                                /* :104-140 */
        double dt_int = 0.0, dt_red = dt;
        while (dt_int < dt) {
            double *c_pointer = d->conc;
            d->get_rhs(d, rhs0);
            for (size_t q = 0; q < n; ++q) c1[q] = d->conc[q] + dt_red * rhs0[q];
            int viol = 0;                                                      /* :121 */
            for (size_t q = 0; q < n; ++q)
                if ((c1[q] - (1.0 + d->relative_change_min) * c_pointer[q]) < 0.0) { viol = 1; break; }
            if (viol && (dt_red > d->dt_min)) {                                /* :126-128 */
                dt_red = dt_red * 0.25;
                d->n_subcycle_warnings++;
                if (d->verbose) fprintf(stderr, " Warning: solver subcycles with dt = %g\n", dt_red);
            } else {
                if (d->adaptive_solver_diagnostics && dt_red < d->last_min_dt) { /* :130-136 */
                    d->last_min_dt = dt_red;
                    minloc_relchange(d, c1, c_pointer, d->last_min_dt_grid_cell);
                }
                for (size_t q = 0; q < n; ++q) d->conc[q] = c1[q];            /* :137-138 */
                dt_int = dt_int + dt_red;
            }
        }
        break;
    }

    case OSED_RK4: {                                                           /* :142-163 */
        double *c_pointer = d->conc;
        rhs1 = dalloc(n, 0.0); rhs2 = dalloc(n, 0.0); rhs3 = dalloc(n, 0.0);
        d->get_rhs(d, rhs0);
        for (size_t q = 0; q < n; ++q) c1[q] = c_pointer[q] + 0.5 * dt * rhs0[q];
        d->conc = c1;
        d->get_rhs(d, rhs1);
        for (size_t q = 0; q < n; ++q) c1[q] = c_pointer[q] + 0.5 * dt * rhs1[q];
        d->get_rhs(d, rhs2);
        for (size_t q = 0; q < n; ++q) c1[q] = c_pointer[q] + dt * rhs2[q];
        d->get_rhs(d, rhs3);
        for (size_t q = 0; q < n; ++q)
            c_pointer[q] = c_pointer[q] + dt * third * (0.5 * rhs0[q] + rhs1[q] + rhs2[q] + 0.5 * rhs3[q]);
        d->conc = c_pointer;
        break;
    }

    case OSED_RK4_38: {                                                        /* :164-185 */
        double *c_pointer = d->conc;
        rhs1 = dalloc(n, 0.0); rhs2 = dalloc(n, 0.0); rhs3 = dalloc(n, 0.0);
        d->get_rhs(d, rhs0);
        for (size_t q = 0; q < n; ++q) c1[q] = c_pointer[q] + third * dt * rhs0[q];
        d->conc = c1;
        d->get_rhs(d, rhs1);
        for (size_t q = 0; q < n; ++q) c1[q] = c_pointer[q] + dt * (rhs1[q] - third * rhs0[q]);
        d->get_rhs(d, rhs2);
        for (size_t q = 0; q < n; ++q) c1[q] = c_pointer[q] + dt * (rhs0[q] - rhs1[q] + rhs2[q]);
        d->get_rhs(d, rhs3);
        for (size_t q = 0; q < n; ++q)
            c_pointer[q] = c_pointer[q] +
                dt * 1.0 / 8.0 * (rhs0[q] + 3.0 * rhs1[q] + 3.0 * rhs2[q] + rhs3[q]);
        d->conc = c_pointer;
        break;
    }
    default: break;
    }
    free(rhs0); free(rhs1); free(rhs2); free(rhs3); free(c1);
}

/* ---- component wrapper ------------------------------------------------------------------ */

/* check_NaN, src/components/fabm_sediment_component.F90:2377-2421 */
int osed_check_nan(const osed_sed *s)
{
    const size_t n3 = N3(s);
    for (size_t c = 0; c < n3; ++c) {
        if (s->base.mask[c] > 0) continue;
        for (int n = 0; n < s->base.nvar; ++n) {
            double v = s->base.conc[c + n3 * n];
            if (v != v) return 1;
        }
    }
    return 0;
}

/* one iteration of the Run loop, src/components/fabm_sediment_component.F90:1715-1732 */
int osed_component_step(osed_sed *s, double dt, int method)
{
    const size_t n3 = N3(s);
    osed_ode_solver(&s->base, dt, method);
    if (osed_check_nan(s)) return 1;
    for (int n = 0; n < s->base.nvar; ++n)                              /* :1726-1732 */
        for (size_t c = 0; c < n3; ++c)
            if (s->base.conc[c + n3 * n] < s->p.minimum[n]) s->base.conc[c + n3 * n] = s->p.minimum[n];
    return 0;
}

/* get_boundary_conditions, src/components/fabm_sediment_component.F90:1865-2030 */
void osed_get_boundary_conditions(osed_sed *s, const double *temperature,
                                  const double *const *csurf, const double *const *wz)
{
    const size_t n2 = N2(s), n3 = N3(s);
    if (temperature)                                                    /* :1930-1935 */
        for (size_t c = 0; c < n2; ++c) s->bdys[c] = temperature[c];
    if (!(s->bcup_dissolved_variables > 0)) return;                     /* :1939 */
    for (int n = 0; n < s->base.nvar; ++n) {
        if (!csurf || !csurf[n]) continue;                              /* :1952-1957 */
        if (s->particulate[n]) {                                        /* :1986 */
            for (size_t c = 0; c < n2; ++c) s->fluxes[c + n2 * n] = -csurf[n][c] * wz[n][c];
        } else {
            for (size_t c = 0; c < n2; ++c) s->bdys[c + n2 * (n + 1)] = csurf[n][c]; /* :2002 */
            if (s->bcup_dissolved_variables == 1) {                     /* :2012-2014 */
                for (size_t c = 0; c < n2; ++c)
                    s->fluxes[c + n2 * n] =
                        -(s->base.conc[c + n3 * n] - s->bdys[c + n2 * (n + 1)]) / s->dz[c] *
                        (s->bioturbation + s->diffusivity + s->bdys[c] * 0.035) * s->porosity[c] /
                        86400. / 10000.;
            } else {
                for (size_t c = 0; c < n2; ++c) s->fluxes[c + n2 * n] = 0.0; /* :2020 */
            }
        }
    }
}

/* 1-D pre-simulation, src/components/fabm_sediment_component.F90:557-632 */
void osed_spinup_column(const osed_sed_nml *nml, const osed_omexdia_params *p, int knum,
                        double dzmin, double dt_min, double relative_change_min,
                        const double *bdys1d, const double *fluxes1d, long nsteps,
                        int method, double *conc1d)
{
    osed_sed s;
    osed_init_grid(&s, 1, 1, knum, dzmin);
    osed_initialize(&s, nml, OSED_MODEL_OMEXDIA_P, p, NULL);
    s.base.conc = conc1d;
    osed_check_domain(&s);
    osed_init_concentrations(&s);
    double bd[OSED_NVAR_OMEXDIA + 1], fl[OSED_NVAR_OMEXDIA];
    memcpy(bd, bdys1d, sizeof(bd)); memcpy(fl, fluxes1d, sizeof(fl));
    s.bdys = bd; s.fluxes = fl;
    s.base.dt_min = dt_min; s.base.relative_change_min = relative_change_min;
    s.bcup_dissolved_variables = 2;                                     /* :608-609 */
    s.base.adaptive_solver_diagnostics = 1;                             /* :610 */
    s.bioturbation_profile = 0;                                         /* :611 */
    for (long t = 0; t < nsteps; ++t) osed_ode_solver(&s.base, 3600.0, method); /* :614-618 */
    osed_finalize(&s);
}

/* ---- CPU baseline harness ---------------------------------------------------------------- */

static double wall_seconds(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
#endif
}

double osed_bench_tiled(int inum, int jnum, int knum, double dzmin, const osed_sed_nml *nml,
                        const osed_omexdia_params *p, const int *mask2d, double *conc,
                        const double *bdys, const double *fluxes_in, double dt, int method,
                        int nsteps, double dt_min, double relative_change_min,
                        int bcup_dissolved, int nthreads, long *subcycles)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > jnum) nthreads = jnum;
    /* The reference's temporaries are automatic (stack) arrays: touched once, then reused.  Keep
     * freed heap blocks mapped so the per-call malloc/free here does not page-fault every step,
     * which would understate the CPU baseline. */
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 64 << 20);
    const int nvar = OSED_NVAR_OMEXDIA;
    osed_sed *tiles = (osed_sed *)calloc(nthreads, sizeof(osed_sed));
    double **tconc = (double **)calloc(nthreads, sizeof(double *));
    int *j0 = (int *)calloc(nthreads + 1, sizeof(int));
    for (int t = 0; t <= nthreads; ++t) j0[t] = (int)((long)jnum * t / nthreads);
    int err = 0;
    /* scatter: one tile per thread, like one PET per DE tile (component :350-387) */
    for (int t = 0; t < nthreads; ++t) {
        osed_sed *s = &tiles[t];
        int jl = j0[t + 1] - j0[t];
        size_t n2l = (size_t)inum * jl, n3l = n2l * knum;
        osed_init_grid(s, inum, jl, knum, dzmin);
        int *mask3 = (int *)calloc(n3l ? n3l : 1, sizeof(int));
        if (mask2d)
            for (int k = 0; k < knum; ++k)
                for (size_t c = 0; c < n2l; ++c) mask3[c + n2l * k] = mask2d[c + (size_t)inum * j0[t]];
        osed_initialize(s, nml, OSED_MODEL_OMEXDIA_P, p, mask3);
        free(mask3);
        s->base.dt_min = dt_min; s->base.relative_change_min = relative_change_min;
        s->bcup_dissolved_variables = bcup_dissolved;
        tconc[t] = dalloc(n3l * nvar, 0.0);
        s->base.conc = tconc[t];
        s->bdys = dalloc(n2l * (nvar + 1), 0.0);
        s->fluxes = dalloc(n2l * nvar, 0.0);
        size_t n2 = (size_t)inum * jnum;
        for (int n = 0; n < nvar; ++n)
            for (int k = 0; k < knum; ++k)
                memcpy(tconc[t] + n2l * (k + (size_t)knum * n),
                       conc + (size_t)inum * j0[t] + n2 * (k + (size_t)knum * n), n2l * sizeof(double));
        for (int n = 0; n < nvar + 1; ++n)
            memcpy(s->bdys + n2l * n, bdys + (size_t)inum * j0[t] + n2 * n, n2l * sizeof(double));
        for (int n = 0; n < nvar; ++n)
            memcpy(s->fluxes + n2l * n, fluxes_in + (size_t)inum * j0[t] + n2 * n, n2l * sizeof(double));
        if (osed_check_domain(s)) err = 1;
        s->poc_data[0] = tconc[t]; s->poc_data[1] = tconc[t] + n3l;
    }
    double t0 = wall_seconds();
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
    for (int t = 0; t < nthreads; ++t)
        for (int st = 0; st < nsteps; ++st)
            if (osed_component_step(&tiles[t], dt, method)) {
#pragma omp atomic write
                err = 2;
            }
    double t1 = wall_seconds();
    long sub = 0;
    for (int t = 0; t < nthreads; ++t) {
        osed_sed *s = &tiles[t];
        int jl = j0[t + 1] - j0[t];
        size_t n2l = (size_t)inum * jl, n2 = (size_t)inum * jnum;
        for (int n = 0; n < nvar; ++n)
            for (int k = 0; k < knum; ++k)
                memcpy(conc + (size_t)inum * j0[t] + n2 * (k + (size_t)knum * n),
                       tconc[t] + n2l * (k + (size_t)knum * n), n2l * sizeof(double));
        sub += s->base.n_subcycle_warnings;
        free(s->bdys); free(s->fluxes); free(tconc[t]);
        osed_finalize(s);
    }
    if (subcycles) *subcycles = sub;
    free(tiles); free(tconc); free(j0);
    return err ? -1.0 : (t1 - t0);
}

/* ---- fused-loop CPU variant ------------------------------------ */

/* One attempt of the adaptive-Euler / Euler step for the columns [c0, c0+nb) of a tile, walking each
 * column once: boundary flux, interface fluxes, omexdia_p rates, dC, c1 = c + dt*rhs and the
 * relative-change test in a single pass (the reference needs >= 40 whole-array passes for the same
 * arithmetic, get_rhs :575-717 + ode_solver :104-140).  Same formulas as osed_get_rhs / osed_diff3d,
 * per column instead of per array; columns are the innermost (SIMD) index.  Hot configuration only:
 * bioturbation_profile != 3, BcUp(particulate) = 1, BcUp(dissolved) in {1,2,3}, BcDown = 3.
 * Returns 1 if some cell violates relative_change_min. */
#define FB 8   /* columns per block: one AVX-512 / two AVX2 vectors of doubles */
static int fused_attempt_block(const osed_sed *s, const double *cin, double *cout, size_t c0, int nb,
                               double dt, double fac, double *flux_out)
{
    const int K = s->base.knum;
    const size_t n2 = N2(s), n3 = N3(s);
    const osed_omexdia_params *p = &s->p;
    const double relaxO2 = 0.04, T0 = 288.15;
    const double E_a = 0.1 * log(1.5) * T0 * (T0 + 10.0);
    const double rLabile = p->rLabile / 86400.0, rSemilabile = p->rSemilabile / 86400.0;
    const double rnit = p->rnit / 86400.0, rODUox = p->rODUox / 86400.0, cmax = p->CprodMax / 86400.0;
    double cpart[FB], cdiss[FB], fT[FB], F[8][FB], Fn[8][FB], cc[8][FB], cn[8][FB], porc[FB], porn[FB];
    int viol = 0;
    for (int b = 0; b < nb; ++b) {
        const double temp = s->bdys[c0 + b];
        const double f_T = exp(-4500.0 * (1.0 / (temp + 273.0) - (1.0 / 288.0)));      /* :648 */
        cpart[b] = s->bioturbation * f_T / 86400.0 / 10000.0;                           /* :652 */
        cdiss[b] = (s->diffusivity + temp * 0.035) / 86400.0 / 10000.0;                 /* :682 */
        fT[b] = exp(-E_a * (1.0 / (temp + 273.15) - 1.0 / T0));
    }
    for (int b = nb; b < FB; ++b) { cpart[b] = cdiss[b] = 0.0; fT[b] = 1.0; }
    /* layer 1 and the upper boundary (diff3d :782-789) */
    for (int b = 0; b < FB; ++b) {
        const size_t c = c0 + (b < nb ? b : 0);
        porc[b] = s->porosity[c];
        for (int n = 0; n < 8; ++n) cc[n][b] = cin[c + n3 * n];
        const double Dp = cpart[b] * (1.0 - porc[b]) * s->bioturbation_factor[c];
        const double Dd = Dp + cdiss[b] * porc[b];
        for (int n = 0; n < 8; ++n) {
            const int part = s->particulate[n];
            const int bc = part ? 1 : s->bcup_dissolved_variables;
            double f = 0.0;
            if (bc == 1) f = s->fluxes[c + n2 * n];
            else if (bc == 2) {
                const double C1 = part ? cc[n][b] * porc[b] : cc[n][b];
                f = -(part ? Dp : Dd) * (C1 - s->bdys[c + n2 * (n + 1)]) / s->dz[c];
            }
            F[n][b] = f;
            if (!part && b < nb) flux_out[c + n2 * n] = f;                              /* :692 */
        }
    }
    for (int k = 0; k < K; ++k) {
        const int has_next = k + 1 < K;
        if (has_next) {
#pragma omp simd
            for (int b = 0; b < FB; ++b) {
                const size_t c = c0 + (b < nb ? b : 0);
                porn[b] = s->porosity[c + n2 * (k + 1)];
                const double intf = 0.5 * (porc[b] + porn[b]);
                const double Dp = cpart[b] * (1.0 - intf) * s->bioturbation_factor[c + n2 * (k + 1)];
                const double Dd = Dp + cdiss[b] * intf;
                const double rdzc = 1.0 / s->dzc[c + n2 * k];
                for (int n = 0; n < 8; ++n) cn[n][b] = cin[c + n2 * (k + 1) + n3 * n];
                for (int n = 0; n < 3; ++n) Fn[n][b] = -Dp * (cn[n][b] * porn[b] - cc[n][b] * porc[b]) * rdzc;
                for (int n = 3; n < 8; ++n) Fn[n][b] = -Dd * (cn[n][b] - cc[n][b]) * rdzc;
            }
        } else {
            for (int n = 0; n < 8; ++n)
                for (int b = 0; b < FB; ++b) Fn[n][b] = 0.0;
        }
#pragma omp simd reduction(| : viol)
        for (int b = 0; b < FB; ++b) {
            const size_t c = c0 + (b < nb ? b : 0);
            const double ldetC = cc[0][b], sdetC = cc[1][b], detP = cc[2][b], po4 = cc[3][b];
            const double no3 = cc[4][b], nh3 = cc[5][b], oxy = cc[6][b], odu = cc[7][b];
            const double Oxicminlim = oxy / (oxy + p->ksO2oxic + relaxO2 * (nh3 + odu));
            const double Denitrilim = (1.0 - oxy / (oxy + p->kinO2denit)) * no3 / (no3 + p->ksNO3denit);
            const double Anoxiclim = (1.0 - oxy / (oxy + p->kinO2anox)) * (1.0 - no3 / (no3 + p->kinNO3anox));
            const double Rescale = 1.0 / (Oxicminlim + Denitrilim + Anoxiclim);
            const double CprodL = rLabile * ldetC, CprodS = rSemilabile * sdetC;
            double Cprod = CprodL + CprodS;
            Cprod = Cprod > cmax ? cmax : Cprod;
            const double Nprod = CprodL * p->NCrLdet + CprodS * p->NCrSdet;
            const double radsP = p->PAds * rSemilabile * po4 * (odu > p->PAdsODU ? odu : p->PAdsODU);
            const double Pprod = rLabile * (1.0 - Oxicminlim) * detP;
            const double OxicMin = Cprod * Oxicminlim * Rescale, Denitrific = Cprod * Denitrilim * Rescale;
            const double AnoxicMin = Cprod * Anoxiclim * Rescale;
            const double Nitri = fT[b] * rnit * nh3 * oxy / (oxy + p->ksO2nitri + relaxO2 * (ldetC + odu));
            const double OduOx = fT[b] * rODUox * odu * oxy / (oxy + p->ksO2oduox + relaxO2 * (nh3 + ldetC));
            double r[8];
            r[0] = -fT[b] * CprodL;  r[1] = -fT[b] * CprodS;
            r[2] = fT[b] * (radsP - Pprod);  r[3] = fT[b] * (Pprod - radsP);
            r[4] = -0.8 * Denitrific + Nitri;  r[5] = (Nprod - Nitri) / (1.0 + p->NH3Ads);
            r[6] = -OxicMin - 2.0 * Nitri - OduOx;  r[7] = AnoxicMin - OduOx;
            const double rpd = 1.0 / (porc[b] * s->dz[c + n2 * k]);     /* dC*(1-por)/por == dF/(por*dz) */
            for (int n = 0; n < 8; ++n) {
                const double rhs = (F[n][b] - Fn[n][b]) * rpd + r[n];
                const double c1 = cc[n][b] + dt * rhs;
                viol |= (c1 - fac * cc[n][b]) < 0.0;
                if (b < nb) cout[c + n2 * k + n3 * n] = c1;
                F[n][b] = Fn[n][b];
                cc[n][b] = cn[n][b];
            }
            porc[b] = porn[b];
        }
    }
    return viol;
}

/* Same contract as osed_bench_tiled, fused loop structure: every attempt reads the state once and writes
 * it once; per tile the accept decision is the reference's whole-tile any() (solver_library.F90:121). */
double osed_bench_fused(int inum, int jnum, int knum, double dzmin, const osed_sed_nml *nml,
                        const osed_omexdia_params *p, const int *mask2d, double *conc,
                        const double *bdys, const double *fluxes_in, double dt, int method,
                        int nsteps, double dt_min, double relative_change_min,
                        int bcup_dissolved, int nthreads, long *subcycles)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > jnum) nthreads = jnum;
    if (method != OSED_EULER && method != OSED_ADAPTIVE_EULER) return -1.0;
    if (nml->bioturbation_profile == 3 || nml->distributed_pom_flux) return -1.0;
    const int nvar = OSED_NVAR_OMEXDIA;
    osed_sed *tiles = (osed_sed *)calloc(nthreads, sizeof(osed_sed));
    double **ta = (double **)calloc(nthreads, sizeof(double *)), **tb = (double **)calloc(nthreads, sizeof(double *));
    int *j0 = (int *)calloc(nthreads + 1, sizeof(int));
    long *subs = (long *)calloc(nthreads, sizeof(long));
    for (int t = 0; t <= nthreads; ++t) j0[t] = (int)((long)jnum * t / nthreads);
    const size_t n2 = (size_t)inum * jnum;
    for (int t = 0; t < nthreads; ++t) {
        osed_sed *s = &tiles[t];
        const int jl = j0[t + 1] - j0[t];
        const size_t n2l = (size_t)inum * jl, n3l = n2l * knum;
        osed_init_grid(s, inum, jl, knum, dzmin);
        int *mask3 = (int *)calloc(n3l ? n3l : 1, sizeof(int));
        if (mask2d)
            for (int k = 0; k < knum; ++k)
                for (size_t c = 0; c < n2l; ++c) mask3[c + n2l * k] = mask2d[c + (size_t)inum * j0[t]];
        osed_initialize(s, nml, OSED_MODEL_OMEXDIA_P, p, mask3);
        free(mask3);
        s->bcup_dissolved_variables = bcup_dissolved;
        ta[t] = dalloc(n3l * nvar, 0.0);
        tb[t] = dalloc(n3l * nvar, 0.0);
        s->bdys = dalloc(n2l * (nvar + 1), 0.0);
        s->fluxes = dalloc(n2l * nvar, 0.0);
        for (int n = 0; n < nvar; ++n)
            for (int k = 0; k < knum; ++k)
                memcpy(ta[t] + n2l * (k + (size_t)knum * n),
                       conc + (size_t)inum * j0[t] + n2 * (k + (size_t)knum * n), n2l * sizeof(double));
        memcpy(tb[t], ta[t], n3l * nvar * sizeof(double));   /* masked columns stay put in both buffers */
        for (int n = 0; n < nvar + 1; ++n)
            memcpy(s->bdys + n2l * n, bdys + (size_t)inum * j0[t] + n2 * n, n2l * sizeof(double));
        for (int n = 0; n < nvar; ++n)
            memcpy(s->fluxes + n2l * n, fluxes_in + (size_t)inum * j0[t] + n2 * n, n2l * sizeof(double));
    }
    const double fac = 1.0 + relative_change_min;
    double t0 = wall_seconds();
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
    for (int t = 0; t < nthreads; ++t) {
        osed_sed *s = &tiles[t];
        const size_t n2l = N2(s), n3l = N3(s);
        double *a = ta[t], *b = tb[t];
        for (int st = 0; st < nsteps; ++st) {
            double dt_int = 0.0, dt_red = dt;
            while (dt_int < dt) {
                int viol = 0;
                for (size_t c0 = 0; c0 < n2l; c0 += FB) {
                    int nb = (int)(n2l - c0 < FB ? n2l - c0 : FB), wet = 0;
                    for (int q = 0; q < nb; ++q) wet |= !(s->base.mask[c0 + q] > 0);
                    if (!wet) continue;
                    if (nb == FB) {
                        int all = 1;
                        for (int q = 0; q < nb; ++q) all &= !(s->base.mask[c0 + q] > 0);
                        if (all) { viol |= fused_attempt_block(s, a, b, c0, nb, dt_red, fac, s->fluxes); continue; }
                    }
                    for (int q = 0; q < nb; ++q)    /* ragged or partly masked block: column by column */
                        if (!(s->base.mask[c0 + q] > 0))
                            viol |= fused_attempt_block(s, a, b, c0 + q, 1, dt_red, fac, s->fluxes);
                }
                if (method == OSED_ADAPTIVE_EULER && viol && dt_red > dt_min) {
                    dt_red *= 0.25;
                    subs[t]++;
                } else {
                    double *tmp = a; a = b; b = tmp;
                    dt_int += (method == OSED_ADAPTIVE_EULER) ? dt_red : dt;
                }
            }
            /* check_NaN + clip (component :1718-1732) in the same sweep over the accepted state */
            for (int n = 0; n < nvar; ++n)
                for (size_t q = 0; q < n3l; ++q) {
                    if (s->base.mask[q] > 0) continue;
                    double *v = &a[q + n3l * n];
                    if (*v < s->p.minimum[n]) *v = s->p.minimum[n];
                }
        }
        ta[t] = a; tb[t] = b;
    }
    double t1 = wall_seconds();
    long sub = 0;
    for (int t = 0; t < nthreads; ++t) {
        osed_sed *s = &tiles[t];
        const int jl = j0[t + 1] - j0[t];
        const size_t n2l = (size_t)inum * jl;
        for (int n = 0; n < nvar; ++n)
            for (int k = 0; k < knum; ++k)
                memcpy(conc + (size_t)inum * j0[t] + n2 * (k + (size_t)knum * n),
                       ta[t] + n2l * (k + (size_t)knum * n), n2l * sizeof(double));
        sub += subs[t];
        free(s->bdys); free(s->fluxes); free(ta[t]); free(tb[t]);
        osed_finalize(s);
    }
    if (subcycles) *subcycles = sub;
    free(tiles); free(ta); free(tb); free(j0); free(subs);
    return t1 - t0;
}

/* ---- flat handle API for the Python test harness ------------------------------------------ */

typedef struct {
    osed_sed s;
    double *conc, *bdys, *fluxes;
} osedpy;

void *osedpy_create(int inum, int jnum, int knum, double dzmin, const osed_sed_nml *nml, int model,
                    const osed_omexdia_params *p, const int *mask2d)
{
    osedpy *h = (osedpy *)calloc(1, sizeof(osedpy));
    if (osed_init_grid(&h->s, inum, jnum, knum, dzmin)) { free(h); return NULL; }
    size_t n2 = (size_t)inum * jnum, n3 = n2 * knum;
    int *mask3 = (int *)calloc(n3 ? n3 : 1, sizeof(int));
    if (mask2d)   /* sed%mask(i,j,:) = (gridmask(i,j).le.0), component :497-501 */
        for (int k = 0; k < knum; ++k)
            for (size_t c = 0; c < n2; ++c) mask3[c + n2 * k] = mask2d[c] > 0;
    osed_initialize(&h->s, nml, model, p, mask3);
    free(mask3);
    h->conc = dalloc(n3 * h->s.base.nvar, 0.0);   /* conc = 0.0_rk, component :531 */
    h->bdys = dalloc(n2 * (h->s.base.nvar + 1), 0.0);
    h->fluxes = dalloc(n2 * h->s.base.nvar, 0.0);
    h->s.base.conc = h->conc;
    h->s.bdys = h->bdys;
    h->s.fluxes = h->fluxes;
    h->s.poc_data[0] = h->conc;
    h->s.poc_data[1] = h->conc + n3;
    return h;
}

void osedpy_destroy(void *hh)
{
    osedpy *h = (osedpy *)hh;
    if (!h) return;
    free(h->conc); free(h->bdys); free(h->fluxes);
    osed_finalize(&h->s);
    free(h);
}

osed_sed *osedpy_sed(void *hh) { return &((osedpy *)hh)->s; }

double *osedpy_ptr(void *hh, int which)
{
    osedpy *h = (osedpy *)hh;
    switch (which) {
    case OSEDPY_CONC: return h->conc;
    case OSEDPY_BDYS: return h->bdys;
    case OSEDPY_FLUXES: return h->fluxes;
    case OSEDPY_POROSITY: return h->s.porosity;
    case OSEDPY_INTF_POROSITY: return h->s.intf_porosity;
    case OSEDPY_BIOTURBATION_FACTOR: return h->s.bioturbation_factor;
    case OSEDPY_PAR: return h->s.par;
    case OSEDPY_PAR_SURFACE: return h->s.par_surface;
    case OSEDPY_TEMP3D: return h->s.temp3d;
    case OSEDPY_FLUX_CAP: return h->s.flux_cap;
    case OSEDPY_BIOMASS: return h->s.biomass;
    case OSEDPY_WEIGHTED_TOC: return h->s.weighted_toc;
    case OSEDPY_DENIT: return h->s.diag_denit;
    case OSEDPY_ZI: return h->s.zi;
    case OSEDPY_ZC: return h->s.zc;
    case OSEDPY_DZ: return h->s.dz;
    case OSEDPY_DZC: return h->s.dzc;
    case OSEDPY_TRANSPORT: return h->s.transport;
    default: return NULL;
    }
}

void osedpy_set_solver(void *hh, double dt_min, double relative_change_min, int bcup_dissolved,
                       int diagnostics, int verbose)
{
    osed_sed *s = osedpy_sed(hh);
    s->base.dt_min = dt_min;
    s->base.relative_change_min = relative_change_min;
    s->bcup_dissolved_variables = bcup_dissolved;
    s->base.adaptive_solver_diagnostics = diagnostics;
    s->base.verbose = verbose;
}

void osedpy_get_solver_diag(void *hh, double *last_min_dt, int cell[4], long *subcycles,
                            double *bioturbation)
{
    osed_sed *s = osedpy_sed(hh);
    *last_min_dt = s->base.last_min_dt;
    for (int q = 0; q < 4; ++q) cell[q] = s->base.last_min_dt_grid_cell[q];
    *subcycles = s->base.n_subcycle_warnings;
    *bioturbation = s->bioturbation;
}

/* src/test/test_Solver.F90:30-44 */
static void test_solver_rhs(osed_rhs_driver *d, double *rhs)
{
    size_t n3 = (size_t)d->inum * d->jnum * d->knum;
    for (size_t q = 0; q < n3 * d->nvar; ++q) rhs[q] = 0.0;
    for (int k = 1; k <= d->knum; ++k)
        for (int j = 1; j <= d->jnum; ++j)
            for (int i = 1; i <= d->inum; ++i)
                for (int v = 0; v < d->nvar; ++v)
                    rhs[(size_t)(i - 1) + (size_t)d->inum * ((j - 1) + (size_t)d->jnum * (k - 1)) + n3 * v] =
                        (i + j + k) * 1.0e-8;
}

void osedpy_test_solver(int inum, int jnum, int knum, int nvar, double *conc, double dt, int method,
                        long nsteps)
{
    osed_rhs_driver d;
    memset(&d, 0, sizeof(d));
    d.inum = inum; d.jnum = jnum; d.knum = knum; d.nvar = nvar;
    d.dt_min = 1.e-9; d.relative_change_min = -0.9;
    d.conc = conc;
    d.last_min_dt = (double)1.e20f;
    d.get_rhs = test_solver_rhs;
    for (long t = 0; t < nsteps; ++t) osed_ode_solver(&d, dt, method);   /* test_Solver.F90:80-82 */
}

/* ---- pelagic <-> soil couplers ------------------------------------------- */

/* pelagic_benthic_coupler Run, src/mediators/pelagic_benthic_coupler.F90:330-480.
 * in[]: 0 oxygen 1 detN 2 detN_wz 3 detC 4 detP 5 detP_wz 6 nitrate 7 ammonium 8 DIN 9 DIP (NULL = absent)
 * csurf(n2,8), wz(n2,3) out.  The oxygen/odu split follows the per-cell intent of :344-349 (the
 * reference assigns the whole arrays inside the i,j loop). */
void osed_pelagic_benthic_coupler_ex(size_t n2, const double *const in[10], double *csurf, double *wz,
                                     int oxy_last_cell)
{
    const double NC_fdet = 0.20, NC_sdet = 0.04, sinking_factor = 0.3;   /* :298,:301-302 */
    for (size_t c = 0; c < n2; ++c) {
        /* oxy_last_cell: the reference assigns the WHOLE oxy/odu arrays inside its i,j loop (:344-349), which
         * leaves every cell with the value of the last one */
        double o2 = in[0][oxy_last_cell ? n2 - 1 : c], detN = in[1][c], vN = in[2][c];
        double CN_det = in[3] ? in[3][c] / detN : 106.0 / 16.0;                     /* :379-394 */
        double fac_fdet = (1.0 - NC_sdet * CN_det) / (NC_fdet - NC_sdet);           /* :395 */
        double fac_sdet = (1.0 - NC_fdet * CN_det) / (NC_sdet - NC_fdet);           /* :396 */
        double din = in[8] ? in[8][c] : 0.0;
        csurf[c + n2 * 0] = fac_fdet * detN;                                        /* :400 */
        csurf[c + n2 * 1] = fac_sdet * detN;                                        /* :403 */
        csurf[c + n2 * 2] = in[4] ? in[4][c] : 1.0 / 16.0 * detN;                   /* :416-420 */
        csurf[c + n2 * 3] = in[9] ? in[9][c] : 1.0 / 16.0 * din;                    /* :466-480 */
        csurf[c + n2 * 4] = in[6] ? in[6][c] : 0.5 * din;                           /* :455-459 */
        csurf[c + n2 * 5] = in[7] ? in[7][c] : 0.5 * din;                           /* :446-450 */
        csurf[c + n2 * 6] = fmax(0.0, o2);                                          /* :346 */
        csurf[c + n2 * 7] = fmax(0.0, -o2);                                         /* :347 */
        wz[c + n2 * 0] = sinking_factor * vN;                                       /* :406 */
        wz[c + n2 * 1] = sinking_factor * vN;                                       /* :408 */
        wz[c + n2 * 2] = sinking_factor * (in[5] ? in[5][c] : vN);                  /* :427-431 */
    }
}

void osed_pelagic_benthic_coupler(size_t n2, const double *const in[10], double *csurf, double *wz)
{
    osed_pelagic_benthic_coupler_ex(n2, in, csurf, wz, 0);
}

/* benthic_pelagic_coupler Run, src/mediators/benthic_pelagic_coupler.F90:211-282.
 * up(n2,8) = <var>_upward_flux_at_soil_surface; out(n2,8): nitrate ammonium DIN DIP detN detC detP oxygen */
void osed_benthic_pelagic_coupler(size_t n2, const double *up, double dinflux_const, double dipflux_const,
                                  double convertN, double NC_fdet, double NC_sdet, double *out)
{
    if (dipflux_const < 0.0) dipflux_const = dinflux_const / 16.0;
    for (size_t c = 0; c < n2; ++c) {
        const double ldetC = up[c], sdetC = up[c + n2], detP = up[c + n2 * 2], po4 = up[c + n2 * 3];
        const double no3 = up[c + n2 * 4], nh3 = up[c + n2 * 5], oxy = up[c + n2 * 6], odu = up[c + n2 * 7];
        out[c + n2 * 0] = convertN * (no3 + dinflux_const / 86400. / 365.);         /* :220 */
        out[c + n2 * 1] = convertN * nh3;                                           /* :222 */
        out[c + n2 * 2] = (no3 + nh3) + dinflux_const / (86400.0 * 365.0);          /* :232-234 */
        out[c + n2 * 3] = po4 + dipflux_const / (86400.0 * 365.0);                  /* :245 */
        out[c + n2 * 4] = convertN * (NC_fdet * ldetC + NC_sdet * sdetC);           /* :258 */
        out[c + n2 * 5] = ldetC + sdetC;                                            /* :264 */
        out[c + n2 * 6] = detP;                                                     /* :274 */
        out[c + n2 * 7] = oxy - odu;                                                /* :281 */
    }
}

/* soil_pelagic_connector Run, src/mediators/soil_pelagic_connector.F90:179-981.
 * up(n2,8) = <var>_upward_flux_at_soil_surface; out(n2,9): nitrate ammonium DIN DIP oxygen odu detN detC detP.
 * want_oxygen / want_odu: which of the two fields the export state holds (branches :660-720). */
void osed_soil_pelagic_connector(size_t n2, const double *up, double dinflux_const, double dipflux_const,
                                 double convertN, double convertP, int want_oxygen, int want_odu, double *out)
{
    if (dipflux_const < 0.0) dipflux_const = dinflux_const / 16.0;                  /* :156 */
    const double year = (double)(86400.0f * 365.0f);                                /* default-real product */
    for (size_t c = 0; c < n2; ++c) {
        const double ldetC = up[c], sdetC = up[c + n2], po4 = up[c + n2 * 3];
        const double no3 = up[c + n2 * 4], nh3 = up[c + n2 * 5], oxy = up[c + n2 * 6], odu = up[c + n2 * 7];
        out[c + n2 * 0] = no3;                                                      /* :333-359 */
        out[c + n2 * 1] = convertN * nh3;                                           /* :409-411 */
        out[c + n2 * 2] = (nh3 + no3 + dinflux_const / year) * convertN;            /* :467-472 */
        out[c + n2 * 3] = convertP * (po4 + dipflux_const / year);                  /* :529-533 */
        if (want_odu && want_oxygen) {                                              /* :660-679 */
            out[c + n2 * 5] = odu;
            out[c + n2 * 4] = oxy;
        } else if (want_odu) {                                                      /* :696-698 */
            out[c + n2 * 5] = odu - oxy;
            out[c + n2 * 4] = oxy;   /* not in the export state; left as the plain flux */
        } else {                                                                    /* :718-720 */
            out[c + n2 * 4] = oxy - odu;
            out[c + n2 * 5] = odu;   /* not in the export state */
        }
        double detN = 0.0, detC = 0.0, detP = 0.0;                                  /* :771,:842,:918 */
        /* omexdia_p exports detritus_{labile,semilabile}_carbon only: no 'detritus*nitrogen' (:764) and no
         * 'detritus*phosphorous' (:911; the model spells it 'phosphorus') field exists to be summed */
        detC = detC + ldetC;                                                        /* :869-874 */
        detC = detC + sdetC;
        out[c + n2 * 6] = detN;
        out[c + n2 * 7] = detC;
        out[c + n2 * 8] = detP;
    }
}

/* pelagic_soil_connector Run, src/mediators/pelagic_soil_connector.F90:176-2122 (2-D export fields, bottom layer
 * of the pelagic fields).  in[]: 0 oxygen 1 odu 2 detN 3 detN_z_velocity 4 detC 5 detP 6 detP_z_velocity 7 nitrate
 * 8 ammonium 9 DIN 10 DIP 11 water_depth 12 tke (NULL = absent).  par[]: sinking_factor sinking_factor_min NC_ldet
 * NC_sdet half_sedimentation_depth half_sedimentation_tke critical_detritus convertN convertP.
 * csurf(n2,8) in the sediment's variable order, wz(n2,3); rows the connector does not write are left alone
 * (oxygen/odu without either import field; the carbon velocities in head_compat mode).
 * head_compat = 0: what the file is written to compute (each export field its own expression, the velocity
 * fields sinking_factor*fac_env*velocity); 1: what the HEAD revision computes where that is defined -- the
 * "velocity" blocks fetch fieldList(1), the CONCENTRATION field, a second time and overwrite it with
 * sinking_factor*fac_env*detN (:1291-1295, :1351-1355); the phosphorus velocity gets the same expression
 * (:1595-1597) unless a detP velocity is imported (:1626-1630); phosphate is recomputed from DIN although DIP
 * was imported (:2092-2094). */
void osed_pelagic_soil_connector(size_t n2, const double *const in[13], const double par[9], int head_compat,
                                 double *csurf, double *wz)
{
    const double sinking_factor = par[0], sinking_factor_min = par[1], NC_ldet = par[2], NC_sdet = par[3];
    const double hsd = par[4], hst = par[5], crit = par[6], convertN = par[7], convertP = par[8];
    const double *oxy = in[0], *odu = in[1], *detN = in[2], *vdetN = in[3], *detC = in[4], *detP = in[5];
    const double *vdetP = in[6], *nit = in[7], *amm = in[8], *din = in[9], *dip = in[10], *depth = in[11], *tke = in[12];
    const int hasN = nit != NULL, hasA = amm != NULL, hasD = din != NULL;
    for (size_t c = 0; c < n2; ++c) {
        double CN_det = 106.0 / 16.0;                                                   /* :1022 */
        if (detC) CN_det = detC[c] / (1E-5f + detN[c]);                                 /* :1063 */
        double fac_ldet = (1.0 - NC_sdet * CN_det) / (NC_ldet - NC_sdet);               /* :1082 */
        if (fac_ldet > CN_det) fac_ldet = CN_det;                                       /* :1084-1087 */
        if (fac_ldet < 0.0) fac_ldet = 0.0;                                             /* :1088-1091 */
        double fac_sdet = CN_det - fac_ldet;                                            /* :1095 */
        double fac_env = 1.0;                                                           /* :1097 */
        if (depth && hsd > 1E-3f)                                                       /* :1150-1155 */
            fac_env = fac_env * (depth[c] * depth[c]) / (depth[c] * depth[c] + hsd * hsd);
        if (tke && hst < 9E9f) fac_env = fac_env * hst / (tke[c] + hst);                /* :1189-1191 */
        fac_env = fac_env + sinking_factor_min / sinking_factor;                        /* :1204 */
        if (detC && crit > 1E-3f && crit < 9E9f) {                                      /* :1215-1227 */
            double x = detC[c] / crit, x2 = x * x;
            fac_env = fac_env * 1.0 / (1.0 + x2 * x2);
        }
        if (!head_compat) {
            csurf[c + n2 * 0] = fac_ldet * convertN * detN[c];                          /* :1282-1284 */
            csurf[c + n2 * 1] = fac_sdet * convertN * detN[c];                          /* :1342-1344 */
            wz[c + n2 * 0] = sinking_factor * fac_env * vdetN[c];
            wz[c + n2 * 1] = sinking_factor * fac_env * vdetN[c];
            wz[c + n2 * 2] = sinking_factor * fac_env * (vdetP ? vdetP[c] : vdetN[c]);
        } else {
            csurf[c + n2 * 0] = sinking_factor * fac_env * detN[c];                     /* :1293-1295 */
            csurf[c + n2 * 1] = sinking_factor * fac_env * detN[c];                     /* :1353-1355 */
            wz[c + n2 * 2] = sinking_factor * fac_env * (vdetP ? vdetP[c] : detN[c]);   /* :1595-1597, :1628-1630 */
        }
        csurf[c + n2 * 2] = detP ? detP[c] : 1.0 / 16.0 * convertN * detN[c];           /* :1521-1523, :1557-1559 */
        double a_out, n_out;
        if (hasA) a_out = convertN * amm[c];                                            /* :1816-1818 */
        else if (hasD && hasN) a_out = convertN * (din[c] - nit[c]);                    /* :1821-1823 */
        else if (hasD) a_out = convertN * 0.5 * din[c];                                 /* :1832-1834 */
        else a_out = convertN * nit[c];                                                 /* :1843-1845 */
        if (hasN) n_out = convertN * nit[c];                                            /* :1930-1932 */
        else if (hasA && hasD) n_out = convertN * (din[c] - amm[c]);                    /* :1939-1943 */
        else if (hasD) n_out = convertN * 0.5 * din[c];                                 /* :1952-1954 */
        else n_out = convertN * amm[c];                                                 /* :1963-1965 */
        csurf[c + n2 * 4] = n_out;
        csurf[c + n2 * 5] = a_out;
        double din_eff;                                                                 /* :2040-2070 */
        if (hasD) din_eff = din[c];
        else if (hasA && hasN) din_eff = nit[c] + amm[c];
        else if (hasA) din_eff = 2 * amm[c];
        else din_eff = 2 * nit[c];
        if (dip && !(head_compat && (hasD || hasA || hasN))) csurf[c + n2 * 3] = convertP * dip[c];
        else csurf[c + n2 * 3] = convertP * (1.0 / 16.0 * convertN * din_eff);          /* :2092-2094 */
        if (oxy && odu) {                                                               /* :806-807 */
            csurf[c + n2 * 6] = oxy[c];
            csurf[c + n2 * 7] = odu[c];
        } else if (odu) {                                                               /* :827-828 */
            csurf[c + n2 * 6] = fmax(0.0, -odu[c]);
            csurf[c + n2 * 7] = fmax(0.0, odu[c]);
        } else if (oxy) {                                                               /* intent of :849-850 */
            csurf[c + n2 * 6] = fmax(0.0, oxy[c]);
            csurf[c + n2 * 7] = fmax(0.0, -oxy[c]);
        }
    }
}

