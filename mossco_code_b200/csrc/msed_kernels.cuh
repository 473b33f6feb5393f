// msed_kernels.cuh -- sm_100a device code of the fabm_sediment column solver.
//
// Shared types (control block, kernel parameters), the step-loop controller and the small
// thread-per-column helper kernels.  The fused RHS + integrator kernel is in msed_column.cuh:
// diff3d transport + omexdia_p reactions (fabm_sediment_driver.F90:575-717,739-825), the
// integrator update (solver_library.F90:99-185), the adaptive-step violation test (:121),
// check_NaN and the minimum clip (fabm_sediment_component.F90:1718-1732) in one pass: per
// ode_solver attempt the state is read once and written once.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include <type_traits>

#include "../../include/msed.h"

namespace msed {

constexpr int NV = MSED_NVAR;
constexpr int MAXK = MSED_MAX_LAYERS;
constexpr int NPART = 3;  // ldetC, sdetC, detP are particulate (main.F90:92-101)
// tunables of the column kernel (overridable with -D for the sweeps in tools/tune_sweep.sh)
#ifndef MSED_COL_BLOCK
#define MSED_COL_BLOCK 128
#endif
#ifndef MSED_COL_MIN_BLOCKS
#define MSED_COL_MIN_BLOCKS 4
#endif
#ifndef MSED_RING_STAGES
#define MSED_RING_STAGES 4
#endif
constexpr int COL_BLOCK = MSED_COL_BLOCK;            // threads (= columns) per CTA of the column kernel
constexpr int COL_MIN_BLOCKS = MSED_COL_MIN_BLOCKS;  // 4 CTAs/SM -> <=128 registers/thread, 16 warps/SM

// integrator stage executed by the column kernel
enum Op : int {
    OP_RHS = 0,       // get_rhs only
    OP_EULER,         // solver_library.F90:99-102
    OP_ADAPTIVE,      // one attempt of :104-140
    OP_RK4_S1, OP_RK4_S2, OP_RK4_S3, OP_RK4_S4,         // :142-163
    OP_RK38_S1, OP_RK38_S2, OP_RK38_S3, OP_RK38_S4      // :164-185
};

// device-resident control block of the step loop (one per handle)
struct Ctl {
    double dt;          // requested ode_solver dt
    double dt_int;      // integrated time inside the current ode_solver call (:106)
    double dt_red;      // current reduced sub-step (:107,:127)
    double dt_min;      // type_rhs_driver%dt_min
    double last_min_dt; // :44
    long long steps_done, steps_target;
    long long rhs_evals, subcycles;
    int cur;            // which of buf[0..1] is sed%conc
    int flags[4];       // [0] relative-change violation (:121)  [1] NaN (component :2392);
                        // [2],[3] the same for the second step of a fused pair (msed_pair.cuh)
    int pairs_disabled; // a fused pair could not be committed: fall back to single steps
    int pair_failures;
    int nan_detected;
    int stop;
    int do_clip;        // component wrapper (check_NaN + clip) on/off
    int diagnostics;    // adaptive_solver_diagnostics
    int minloc_request; // set when last_min_dt decreased (:131-135)
    int step_completed; // 1 if the last controller invocation finished an ode_solver call
};

struct OmexDev {  // hzg_omexdia_p parameters, rates already per second
    double rLabile, rSemilabile, NCrLdet, NCrSdet, PAds_rS, PAdsODU, rNH3Ads, CprodMax;
    double rnit, ksO2nitri, rODUox, ksO2oduox, ksO2oxic, ksNO3denit, kinO2denit;
    double kinNO3anox, kinO2anox, E_a;
    double minimum[NV];
};

struct KParams {
    double *buf[2];          // ping-pong state, [nvar][K][ld]
    double *aux1, *aux2;     // RK accumulators
    double *rhs_out;         // OP_RHS target
    const double *por;       // [K][ld]
    const double *bdys;      // [nvar+1][ld]
    double *fluxes;          // [nvar][ld]
    const unsigned char *mask;  // [ncol]
    Ctl *ctl;
    size_t ld;
    int ncol, K, inum;
    int col0, col_end;       // column range [col0, col_end) of this launch (chunked launches overlap PCIe)
    int i_offset, j_offset;
    int bcup_diss, bcup_part, profile;
    int use_ctl;             // 0: OP_RHS / plain launch with p.dt and buffer 0
    int por_mode;            // 0: 3-D porosity field, 1: portab[k], 2: por(:,:,1)*portab[k]
    double dt;
    double fac;              // 1 + relative_change_min
    double bioturbation, diffusivity;
    double pom_flux_rate;    // pom_flux_max/86400
    double beta, b, L1, L2, poc_factor[2], cumdepth_last;
    OmexDev om;
    double dz[MAXK], rdzc[MAXK], bf[MAXK], e1[MAXK], e2[MAXK], portab[MAXK];
    double *denit_out;       // [K][ld]: FABM denit diagnostic of the second step of a call's last pair, or null
    const int *colmap;       // pair_kernel on a masked tile: indices of the wet columns, ascending; col0/col_end
                             // then count wet columns (null: identity)
    const double *in_ovr;    // pair_kernel<.., OVR>: explicit input / output state buffers of a launch inside a
    double *out_ovr;         // chunk-major sequence (msed.cu run_steps), instead of buf[cur] / buf[1-cur]
};

// loaders, reaction term and the fused column kernel
#include "msed_column.cuh"
// two Euler / adaptive-Euler steps per pass over HBM (speculative, rollback-free)
#include "msed_pair.cuh"
// a chain of steps with the column in registers: warp per column, lane per layer (knum <= 32)
#include "msed_chain.cuh"
#include "msed_rkpair.cuh"

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
            return;
        }
        if (c->diagnostics && c->dt_red < c->last_min_dt) {  // :130-135
            c->last_min_dt = c->dt_red;
            c->minloc_request = 1;
        }
        c->cur ^= 1;                      // conc = c1, :137
        c->dt_int = c->dt_int + c->dt_red;  // :138
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
// rows of a staging area [..][ld_out]; the per-cell intent of the oxygen/odu split is implemented
// (the reference assigns the whole array inside its i,j loop, :344-349).
struct P2BIn {
    const double *oxygen, *detN, *detN_wz, *detC, *detP, *detP_wz, *nitrate, *ammonium, *DIN, *DIP;
};
__global__ void pelagic_benthic_kernel(double *csurf, double *wz, P2BIn in, size_t ld_out, int ncol)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const double NC_fdet = 0.20, NC_sdet = 0.04, sinking_factor = 0.3;   // :298,:301-302
    const double o2 = in.oxygen[col];
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
__global__ void negate_rows_kernel(double *dst, const double *src, size_t ld, int ncol, int rows)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    for (int q = 0; q < rows; ++q) dst[(size_t)q * ld + col] = -src[(size_t)q * ld + col];
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
            if (v < minimum.v[n]) s[q] = minimum.v[n];
        }
    if (nanf) { c->nan_detected = 1; c->stop = 1; }
}

}  // namespace msed
