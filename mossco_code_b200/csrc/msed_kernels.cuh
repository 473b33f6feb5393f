// msed_kernels.cuh -- sm_100a device code of the fabm_sediment column solver that lives in msed.cu:
// the step-loop controllers and the small thread-per-column helper kernels.
//
// The fused RHS + integrator kernels are in msed_column.cuh (one step per launch), msed_pair.cuh (two
// sub-steps per launch), msed_chain.cuh / msed_strip.cuh (many sub-steps per launch with the column in
// registers) and msed_rkpair.cuh (two Runge-Kutta stages per launch): diff3d transport + omexdia_p reactions
// (fabm_sediment_driver.F90:575-717,739-825), the integrator update (solver_library.F90:99-185), the
// adaptive-step violation test (:121), check_NaN and the minimum clip (fabm_sediment_component.F90:1718-1732)
// in one pass.
#pragma once

#include "msed_types.cuh"

namespace msed {

// loaders and the reaction term (the helper kernels below evaluate diagnostics with the same inline code);
// the stepping kernels themselves are instantiated in the msed_tu_*.cu translation units (msed_launch.h)
#include "msed_column.cuh"

// ---------------------------------------------------------------------------------------------
// step-loop controller: the scalar part of ode_solver (:108,:126-139) and of the component
// loop, executed on device so the host never has to synchronise inside a Run
// ---------------------------------------------------------------------------------------------
__global__ void controller_kernel(Ctl *c, int method)
{
    c->step_completed = 0;
    if (c->stop || c->steps_done >= c->steps_target) return;
    const int viol = c->flags[0], nanf = c->flags[1];
    c->flags[0] = 0;
    c->flags[1] = 0;
    if (method == MSED_ADAPTIVE_EULER) {
        c->rhs_evals += 1;
        if (viol && c->dt_red > c->dt_min) {  // :126-128
            c->dt_red = c->dt_red * 0.25;
            c->subcycles += 1;
            if (c->step_accepts == 0) c->step_rej_first += 1;
            else c->step_rej_later += 1;
            return;
        }
        if (c->diagnostics && c->dt_red < c->last_min_dt) {  // :130-135
            c->last_min_dt = c->dt_red;
            c->minloc_request = 1;
        }
        c->cur ^= 1;                      // conc = c1, :137
        c->dt_int = c->dt_int + c->dt_red;  // :138
        c->step_accepts += 1;
        if (c->dt_int < c->dt) return;    // :108
    } else if (method == MSED_EULER) {
        c->rhs_evals += 1;
        c->cur ^= 1;
    } else {
        c->rhs_evals += 4;
    }
    if (nanf && c->do_clip) {
        c->nan_detected = 1;
        c->stop = 1;
    }
    c->steps_done += 1;
    c->step_completed = 1;
    c->dt_int = 0.0;
    c->dt_red = c->dt;
    c->last_depth = c->step_rej_first;
    c->last_irregular = c->step_rej_later > 0;
    c->step_rej_first = c->step_rej_later = c->step_accepts = 0;
}

// Commits one fused group -- a pair launch, a chain launch, or all pair launches of a chunk-major coupling
// interval -- or disables fused launches so that the host redoes the same attempts one by one from the
// untouched state.  The group was planned on the host (msed.cu run_steps) as a definite piece of the
// reference's attempt sequence (solver_library.F90:104-140): which attempts are rejected, which sub-steps are
// accepted, where ode_solver calls end.  It is committed only if every flag agrees with that plan:
//   - every planned rejection was seen (its flag slot is up: some cell violated relative_change_min at the
//     larger step, :121,:126), otherwise the reference would have accepted the larger step;
//   - no accepted sub-step violates while it could still be rejected (:126), otherwise the reference would
//     have gone to a smaller step;
//   - check_NaN (component :1718) found nothing.
__global__ void plan_controller_kernel(Ctl *c, PlanCommit pc)
{
    c->step_completed = 0;
    if (c->stop || c->pairs_disabled || c->steps_done != pc.gate_steps) return;
    const int own = c->flags[0] | c->flags[2], nanf = c->flags[1] | c->flags[3];
    int up_all = 1, first_missing = -1;
    for (int s = 0; s < pc.up_slots; ++s)
        if (c->flags[FLAG_UP0 + s] == 0) {
            up_all = 0;
            if (first_missing < 0) first_missing = s;
        }
    const int stopped_at = c->flags[FLAG_FAIL] > 0 ? 64 - c->flags[FLAG_FAIL] : -1;   // chain_kernel
    for (int s = 0; s < MSED_NFLAGS; ++s) c->flags[s] = 0;
    if ((pc.own_rejectable && own) || (c->do_clip && nanf) || !up_all) {
        c->pairs_disabled = 1;
        c->pair_failures += 1;
        // a step of the group at or before which the plan went wrong (chains: run_steps re-runs the steps in
        // front of it as a shorter chain): the first planned rejection that was not seen, the step a warp
        // stopped at -- whichever comes first
        int f = -1;
        if (first_missing >= 0 && pc.depth > 0) f = first_missing / pc.depth;
        if (stopped_at >= 0 && (f < 0 || stopped_at < f)) f = stopped_at;
        c->fail_step = f;
        return;
    }
    if (pc.flip) c->cur ^= 1;
    c->steps_done += pc.steps;
    c->fused_steps += pc.steps;
    c->fused_launches += pc.launches;
    c->rhs_evals += pc.rhs_evals;
    c->subcycles += pc.subcycles;
    c->dt_int = pc.dt_int;
    c->dt_red = pc.dt_red;
    c->step_completed = (pc.steps > 0 && pc.dt_int == 0.0) ? 1 : 0;
    if (pc.steps > 0) {
        c->last_depth = pc.depth;
        c->last_irregular = 0;
    }
    if (pc.dt_int == 0.0) {
        c->step_rej_first = c->step_rej_later = c->step_accepts = 0;
    } else {   // the group ends inside an ode_solver call: what the single-step controller needs to carry on
        c->step_rej_first = pc.depth;
        c->step_rej_later = 0;
        c->step_accepts = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// small helper kernels (initialisation / boundary / export); all thread-per-column
// ---------------------------------------------------------------------------------------------

// porosity(:,:,k) = table[k]; masked -> 1 (driver :280,:431)
__global__ void fill_porosity_kernel(double *por, const unsigned char *mask, size_t ld, int ncol, int K,
                                     const double *table)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const bool m = mask[col] != 0;
    for (int k = 0; k < K; ++k) por[(size_t)k * ld + col] = m ? 1.0 : table[k];
}

// update_porosity(from_surface=.true.), driver :409-413,:431
__global__ void porosity_from_surface_kernel(double *por, const double *surf, const unsigned char *mask,
                                             size_t ld, int ncol, int K, double porosity_fac,
                                             const double *zc)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const bool m = mask[col] != 0;
    const double p0 = surf[col];
    por[col] = m ? 1.0 : p0;
    for (int k = 1; k < K; ++k) {
        // explicit IEEE operations (no FMA contraction): porosity must be bit-identical to the
        // Fortran expression at driver :411-412
        const double v = __dmul_rn(p0, __dsub_rn(1.0, __dmul_rn(porosity_fac, __dsub_rn(zc[k], zc[0]))));
        por[(size_t)k * ld + col] = m ? 1.0 : v;
    }
}

// fabm_sed_check_domain porosity conditions on unmasked cells (driver :503-511)
__global__ void check_porosity_kernel(const double *por, const unsigned char *mask, size_t ld, int ncol,
                                      int K, int *flags)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol || mask[col]) return;
    bool le0 = false, gt1 = false;
    for (int k = 0; k < K; ++k) {
        const double v = por[(size_t)k * ld + col];
        le0 |= (v <= 0.0);
        gt1 |= (v > 1.0);
    }
    if (le0) atomicOr(&flags[0], 1);
    if (gt1) atomicOr(&flags[1], 1);
}

// where(mask>0) porosity = 1 ; conc = 1e20  (driver :431,:535,:541)
__global__ void apply_mask_kernel(double *por, double *b0, double *b1, const unsigned char *mask, size_t ld,
                                  int ncol, int K)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol || !mask[col]) return;
    for (int k = 0; k < K; ++k) por[(size_t)k * ld + col] = 1.0;
    for (int q = 0; q < NV * K; ++q) {
        b0[(size_t)q * ld + col] = 1.0e20;
        b1[(size_t)q * ld + col] = 1.0e20;
    }
}

// conc = missing_value on the land columns of ONE state buffer: the staging buffer that a chunk-major Run
// rotates in as state buffer (msed.cu run_steps) was never written there, because every stepping kernel
// skips land columns (driver :464,:535)
__global__ void fill_masked_kernel(double *b, const unsigned char *mask, size_t ld, int ncol, int rows,
                                   double missing)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol || !mask[col]) return;
    for (int q = 0; q < rows; ++q) b[(size_t)q * ld + col] = missing;
}

// init_concentrations, driver :455-468
struct InitVals { double v[NV]; };
__global__ void init_conc_kernel(double *b0, double *b1, const double *por, const unsigned char *mask,
                                 size_t ld, int ncol, int K, InitVals iv, double missing)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const bool m = mask[col] != 0;
    for (int k = 0; k < K; ++k) {
        const double pk = por[(size_t)k * ld + col];
        for (int n = 0; n < NV; ++n) {
            const double v = m ? missing : iv.v[n] / pk;
            b0[(size_t)(n * K + k) * ld + col] = v;
            b1[(size_t)(n * K + k) * ld + col] = v;
        }
    }
}

// sed%conc(i,j,:,:) = sed1d%conc(1,1,:,:) for unmasked columns, component :628-632
__global__ void broadcast_column_kernel(double *b0, double *b1, const double *col1d,
                                        const unsigned char *mask, size_t ld, int ncol, int K)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol || mask[col]) return;
    for (int q = 0; q < NV * K; ++q) {
        const double v = col1d[q];
        b0[(size_t)q * ld + col] = v;
        b1[(size_t)q * ld + col] = v;
    }
}

__global__ void copy_state_kernel(double *dst, const double *src, size_t ld, int ncol, int rows)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    for (int q = 0; q < rows; ++q) dst[(size_t)q * ld + col] = src[(size_t)q * ld + col];
}

// get_boundary_conditions, component :1930-2020.  present bit n set: csurf[n] in the import state
struct BcPtrs { const double *temperature; const double *csurf[NV]; const double *wz[NV]; };
__global__ void boundary_kernel(double *bdys, double *fluxes, const double *conc, const double *por,
                                BcPtrs in, size_t ld, size_t ld_in, int ncol, int K, int bcup_diss,
                                double bioturbation, double diffusivity, double dz0)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    if (in.temperature) bdys[col] = in.temperature[col];            // :1930-1935
    if (!(bcup_diss > 0)) return;                                   // :1939
    const double temp = bdys[col];
    for (int n = 0; n < NV; ++n) {
        if (!in.csurf[n]) continue;
        if (n < NPART) {                                            // :1986
            fluxes[(size_t)n * ld + col] = -in.csurf[n][col] * in.wz[n][col];
        } else {
            const double cs = in.csurf[n][col];
            bdys[(size_t)(n + 1) * ld + col] = cs;                  // :2002
            if (bcup_diss == 1) {                                   // :2012-2014
                const double c1 = conc[(size_t)(n * K) * ld + col];
                fluxes[(size_t)n * ld + col] =
                    __ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(__ddiv_rn(-(c1 - cs), dz0),
                                                             __dadd_rn(bioturbation + diffusivity,
                                                                       __dmul_rn(temp, 0.035))),
                                                  por[col]),
                                        86400.), 10000.);
            } else {
                fluxes[(size_t)n * ld + col] = 0.0;                 // :2020
            }
        }
    }
    (void)ld_in;
}

// pelagic_benthic_coupler Run (src/mediators/pelagic_benthic_coupler.F90:330-480), one thread per
// column.  Writes the 8 surface concentrations into csurf rows and the 3 sinking velocities into wz
// rows of a staging area [..][ld_out]; the per-cell intent of the oxygen/odu split is implemented by default
// (the reference assigns the whole array inside its i,j loop, :344-349: see oxy_last_cell).
struct P2BIn {
    const double *oxygen, *detN, *detN_wz, *detC, *detP, *detP_wz, *nitrate, *ammonium, *DIN, *DIP;
};
__global__ void pelagic_benthic_kernel(double *csurf, double *wz, P2BIn in, size_t ld_out, int ncol,
                                       int oxy_last_cell)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const double NC_fdet = 0.20, NC_sdet = 0.04, sinking_factor = 0.3;   // :298,:301-302
    // MSED_COMPAT_P2B_OXYGEN_LAST_CELL: the reference assigns the whole oxy/odu arrays inside its i,j loop
    // (:344-349), so every column ends up with the value of the tile's last cell
    const double o2 = in.oxygen[oxy_last_cell ? ncol - 1 : col];
    const double detN = in.detN[col];
    const double vN = in.detN_wz[col];
    const double CN = in.detC ? __ddiv_rn(in.detC[col], detN) : 106.0 / 16.0;           // :379-394
    const double fac_f = __ddiv_rn(__dsub_rn(1.0, __dmul_rn(NC_sdet, CN)), NC_fdet - NC_sdet);  // :395
    const double fac_s = __ddiv_rn(__dsub_rn(1.0, __dmul_rn(NC_fdet, CN)), NC_sdet - NC_fdet);  // :396
    csurf[0 * ld_out + col] = __dmul_rn(fac_f, detN);                                   // :400
    csurf[1 * ld_out + col] = __dmul_rn(fac_s, detN);                                   // :403
    csurf[2 * ld_out + col] = in.detP ? in.detP[col] : __dmul_rn(1.0 / 16.0, detN);     // :416-420
    const double din = in.DIN ? in.DIN[col] : 0.0;
    csurf[3 * ld_out + col] = in.DIP ? in.DIP[col] : __dmul_rn(1.0 / 16.0, din);        // :466-480
    csurf[4 * ld_out + col] = in.nitrate ? in.nitrate[col] : __dmul_rn(0.5, din);       // :455-459
    csurf[5 * ld_out + col] = in.ammonium ? in.ammonium[col] : __dmul_rn(0.5, din);     // :446-450
    csurf[6 * ld_out + col] = o2 > 0.0 ? o2 : 0.0;                                      // max(0, o2)  :346
    csurf[7 * ld_out + col] = -o2 > 0.0 ? -o2 : 0.0;                                    // max(0,-o2)  :347
    wz[0 * ld_out + col] = __dmul_rn(sinking_factor, vN);                               // :406
    wz[1 * ld_out + col] = __dmul_rn(sinking_factor, vN);                               // :408
    wz[2 * ld_out + col] = __dmul_rn(sinking_factor, in.detP_wz ? in.detP_wz[col] : vN);  // :427-431
}

// benthic_pelagic_coupler Run (src/mediators/benthic_pelagic_coupler.F90:211-282).
// out rows: 0 nitrate 1 ammonium 2 DIN 3 DIP 4 detN 5 detC 6 detP 7 oxygen
__global__ void benthic_pelagic_kernel(double *out, const double *fluxes, size_t ld, int ncol,
                                       double dinflux_const, double dipflux_const, double convertN,
                                       double NC_fdet, double NC_sdet)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    double up[NV];
    for (int n = 0; n < NV; ++n) up[n] = -fluxes[(size_t)n * ld + col];                 // component :1819
    const double year = 86400.0 * 365.0;
    out[0 * ld + col] = __dmul_rn(convertN, __dadd_rn(up[4], __ddiv_rn(__ddiv_rn(dinflux_const, 86400.), 365.)));  // :220
    out[1 * ld + col] = __dmul_rn(convertN, up[5]);                                      // :222
    out[2 * ld + col] = __dadd_rn(__dadd_rn(up[4], up[5]), __ddiv_rn(dinflux_const, year));  // :232-234
    out[3 * ld + col] = __dadd_rn(up[3], __ddiv_rn(dipflux_const, year));                // :245
    out[4 * ld + col] = __dmul_rn(convertN, __dadd_rn(__dmul_rn(NC_fdet, up[0]), __dmul_rn(NC_sdet, up[1])));  // :258
    out[5 * ld + col] = __dadd_rn(up[0], up[1]);                                         // :264
    out[6 * ld + col] = up[2];                                                           // :274
    out[7 * ld + col] = __dsub_rn(up[6], up[7]);                                         // :281
}

// soil_pelagic_connector Run (src/mediators/soil_pelagic_connector.F90:179-981), the generic successor of
// benthic_pelagic_coupler.  out rows: 0 nitrate 1 ammonium 2 DIN 3 DIP 4 oxygen 5 odu 6 detC 7 zero.
// Row 7 serves both detritus nitrogen and detritus phosphorus: the connector sums the import fields that
// match 'detritus*nitrogen_upward_flux_at_soil_surface' (:764) and 'detritus*phosphorous_upward_flux_at_
// soil_surface' (:911); omexdia_p exports neither (its P pool is spelled 'detritus_labile_phosphorus'), so
// both sums stay at their initial 0.0 (:771-773, :918-920).
// oxy_mode: 3 = oxygen and odu both wanted (plain copies, :660-679), 2 = only odu (odu - oxygen, :696-698),
// 1 = only oxygen (oxygen - odu, :718-720).
__global__ void soil_pelagic_kernel(double *out, const double *fluxes, size_t ld, int ncol, double din_per_s,
                                    double dip_per_s, double convertN, double convertP, int oxy_mode)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    double up[NV];
    for (int n = 0; n < NV; ++n) up[n] = -fluxes[(size_t)n * ld + col];                 // component :1819
    out[0 * ld + col] = up[4];                                                           // :333-359 (no convertN)
    out[1 * ld + col] = __dmul_rn(convertN, up[5]);                                      // :409-411
    out[2 * ld + col] = __dmul_rn(__dadd_rn(__dadd_rn(up[5], up[4]), din_per_s), convertN);  // :467-472
    out[3 * ld + col] = __dmul_rn(convertP, __dadd_rn(up[3], dip_per_s));                // :529-533
    out[4 * ld + col] = (oxy_mode == 1) ? __dsub_rn(up[6], up[7]) : up[6];               // :677-679, :718-720
    out[5 * ld + col] = (oxy_mode == 2) ? __dsub_rn(up[7], up[6]) : up[7];               // :660-662, :696-698
    out[6 * ld + col] = __dadd_rn(__dadd_rn(0.0, up[0]), up[1]);                         // :842-874
    out[7 * ld + col] = 0.0;                                                             // :771-773, :918-920
}

// pelagic boxes take up the bed flux: conc = conc + bfl*dt/layer_height where layer_height > 0
// (src/components/fabm_pelagic_component.F90:2100-2105); bfl = upward flux = -fluxes (component :1819)
__global__ void pelagic_flux_kernel(double *pel, const double *fluxes, const double *height,
                                    const unsigned char *mask, size_t ld, int ncol, double dtc)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol || mask[col]) return;
    const double hgt = height[col];
    if (!(hgt > 0.0)) return;
    for (int n = 0; n < NV; ++n) {
        const size_t q = (size_t)n * ld + col;
        pel[q] = __dadd_rn(pel[q], __ddiv_rn(__dmul_rn(-fluxes[q], dtc), hgt));
    }
}

// -fluxes(:,:,n) for the export state, component :1819
__global__ void negate_rows_kernel(double *dst, const double *src, size_t ld_dst, size_t ld, int ncol, int rows)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    for (int q = 0; q < rows; ++q) dst[(size_t)q * ld_dst + col] = -src[(size_t)q * ld + col];
}

// derived 3-D export fields (driver :597-604,:434-435,:284-291,:627-644 and the FABM denit diagnostic)
__global__ void field_kernel(double *out, int which, KParams p, const double *par_surface,
                             const double *state, const double *cumdepth, double k_par, double missing_temp)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= p.ncol) return;
    const int K = p.K;
    const size_t ld = p.ld;
    const bool m = p.mask[col] != 0;
    const double temp = p.bdys[col];
    double avg = 0.0;
    if (which == MSED_FIELD_BIOMASS || which == MSED_FIELD_BIOTURBATION) {
        double s = 0.0;
        for (int k = 0; k < K; ++k)
            s += p.dz[k] * wtoc_cell(p, p.por[(size_t)k * ld + col], state[(size_t)k * ld + col],
                                     state[(size_t)(K + k) * ld + col]);
        avg = s / p.cumdepth_last;
    }
    double fT = 1.0;
    if (which == MSED_FIELD_DENIT) fT = exp(-p.om.E_a * (1.0 / (temp + 273.15) - 1.0 / 288.15));
    double cap_prev = 0.0;
    for (int k = 0; k < K; ++k) {
        const size_t q = (size_t)k * ld + col;
        const double pk = p.por[q];
        double v = 0.0;
        switch (which) {
        case MSED_FIELD_TEMPERATURE: v = m ? missing_temp : temp; break;
        case MSED_FIELD_PAR: v = m ? 0.0 : par_surface[col] * exp(-cumdepth[k] / k_par); break;
        case MSED_FIELD_INTF_POROSITY:
            v = (k == 0) ? pk : 0.5 * (p.por[(size_t)(k - 1) * ld + col] + pk);
            break;
        case MSED_FIELD_FLUX_CAP: {
            double cap = p.pom_flux_rate * (1.0 - pk) * p.dz[k];
            if (k + 1 > 2 && cap > cap_prev) cap = cap_prev;
            cap_prev = cap;
            v = cap;
        } break;
        case MSED_FIELD_WEIGHTED_TOC:
            v = m ? 0.0 : wtoc_cell(p, pk, state[q], state[(size_t)(K + k) * ld + col]);
            break;
        case MSED_FIELD_BIOMASS:
        case MSED_FIELD_BIOTURBATION: {
            if (m) { v = (which == MSED_FIELD_BIOMASS) ? 0.0 : 1.0; break; }
            const double wt = wtoc_cell(p, pk, state[q], state[(size_t)(K + k) * ld + col]);
            const double biomass = wt * p.e1[k] * avg / (p.L1 + p.L2 * p.e2[k]);
            v = (which == MSED_FIELD_BIOMASS) ? biomass : p.beta * pow(biomass, p.b) / wt;
        } break;
        case MSED_FIELD_DENIT: {
            if (m) { v = 0.0; break; }
            double c[NV], r[NV], d;
            for (int n = 0; n < NV; ++n) c[n] = state[(size_t)(n * K + k) * ld + col];
            omexdia_rates(p.om, c, fT, r, &d);
            v = d;
        } break;
        default: break;
        }
        out[q] = v;
    }
}

// the denitrification diagnostic stored by the last fused pair of a call -> <name>_in_soil layout
__global__ void denit_copy_kernel(double *out, const double *denit, const unsigned char *mask, size_t ld,
                                  int ncol, int K)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const bool m = mask[col] != 0;
    for (int k = 0; k < K; ++k) out[(size_t)k * ld + col] = m ? 0.0 : denit[(size_t)k * ld + col];
}

// minloc((c1-c)/c) in Fortran array order, NaNs skipped (solver_library.F90:133-134).
// single block; only used with adaptive_solver_diagnostics (1-D spin-up, standalone harness)
__global__ void minloc_kernel(Ctl *c, const double *b0, const double *b1, size_t ld, int ncol, int rows,
                              double *best_val, long long *best_idx)
{
    if (!c->minloc_request) return;
    const double *c1 = c->cur ? b1 : b0;   // accepted state
    const double *c0 = c->cur ? b0 : b1;   // state before the accepted sub-step
    __shared__ double sv[256];
    __shared__ long long si[256];
    double bv = 0.0;
    long long bi = -1;
    const long long total = (long long)rows * ncol;
    for (long long q = threadIdx.x; q < total; q += blockDim.x) {
        const long long row = q / ncol, col = q % ncol;
        const double a = c0[(size_t)row * ld + col], b = c1[(size_t)row * ld + col];
        const double v = (b - a) / a;
        if (v != v) continue;
        if (bi < 0 || v < bv) { bv = v; bi = q; }   // q ascending per thread: first minimum kept
    }
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        long long i = -1;
        for (int t = 0; t < blockDim.x; ++t) {
            if (si[t] < 0) continue;
            if (i < 0 || sv[t] < v || (sv[t] == v && si[t] < i)) { v = sv[t]; i = si[t]; }
        }
        *best_val = v;
        *best_idx = i;
        c->minloc_request = 0;
    }
}

// conc = max(conc, minimum) as a separate pass (only used together with minloc diagnostics)
__global__ void clip_kernel(Ctl *c, double *b0, double *b1, const unsigned char *mask, size_t ld, int ncol,
                            int K, InitVals minimum)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol || mask[col] || !c->step_completed) return;
    double *s = c->cur ? b1 : b0;
    bool nanf = false;
    for (int n = 0; n < NV; ++n)
        for (int k = 0; k < K; ++k) {
            const size_t q = (size_t)(n * K + k) * ld + col;
            const double v = s[q];
            nanf |= (v != v);
            if (__double_as_longlong(v) < __double_as_longlong(minimum.v[n])) s[q] = minimum.v[n];   // clip_min
        }
    if (nanf) { c->nan_detected = 1; c->stop = 1; }
}

// pelagic -> soil connector, whole-domain diagnostics, fp64 micro-benchmark
#include "msed_aux.cuh"

}  // namespace msed
