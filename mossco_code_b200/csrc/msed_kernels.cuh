// msed_kernels.cuh -- sm_100a device code of the fabm_sediment column solver.
//
// One thread owns one sediment column and streams it top -> bottom, holding a two-layer
// register window (layer k and k+1: 8 state variables each) plus the 8 interface fluxes of
// the upper interface.  Lanes of a warp own adjacent columns, so every load/store of the
// [nvar][knum][ld] state is a fully coalesced 256-byte warp access.  The RHS
// (diff3d transport + omexdia_p reactions, fabm_sediment_driver.F90:575-717,739-825), the
// integrator update (solver_library.F90:99-185), the adaptive-step violation test (:121),
// check_NaN and the minimum clip (fabm_sediment_component.F90:1718-1732) are fused: per
// ode_solver attempt the state is read once and written once.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/msed.h"

namespace msed {

constexpr int NV = MSED_NVAR;
constexpr int MAXK = MSED_MAX_LAYERS;
constexpr int NPART = 3;  // ldetC, sdetC, detP are particulate (main.F90:92-101)
constexpr int COL_BLOCK = 128;      // threads (= columns) per CTA of the column kernel
constexpr int COL_MIN_BLOCKS = 4;   // 4 CTAs/SM -> <=128 registers/thread, 16 warps/SM

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
    int flags[2];       // [0] relative-change violation (:121)  [1] NaN (component :2392)
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
    int i_offset, j_offset;
    int bcup_diss, bcup_part, profile;
    int use_ctl;             // 0: OP_RHS / plain launch with p.dt and buffer 0
    double dt;
    double fac;              // 1 + relative_change_min
    double bioturbation, diffusivity;
    double pom_flux_rate;    // pom_flux_max/86400
    double beta, b, L1, L2, poc_factor[2], cumdepth_last;
    OmexDev om;
    double dz[MAXK], rdzc[MAXK], bf[MAXK], e1[MAXK], e2[MAXK];
};

__device__ __forceinline__ double ld_state(const double *p) { return *p; }
__device__ __forceinline__ double ld_ro(const double *p) { return __ldg(p); }

// hzg_omexdia_p local rates for one cell: SURVEY.md Appendix B (frozen project spec; the FABM
// source is not part of the reference tree).  fT is the per-column Arrhenius factor.
__device__ __forceinline__ void omexdia_rates(const OmexDev &m, const double (&c)[NV], double fT,
                                              double (&r)[NV], double *denit)
{
    const double ldetC = c[0], sdetC = c[1], detP = c[2], po4 = c[3];
    const double no3 = c[4], nh3 = c[5], oxy = c[6], odu = c[7];
    const double relaxO2 = 0.04;

    const double Oxicminlim = oxy / (oxy + m.ksO2oxic + relaxO2 * (nh3 + odu));
    const double Denitrilim = (1.0 - oxy / (oxy + m.kinO2denit)) * no3 / (no3 + m.ksNO3denit);
    const double Anoxiclim = (1.0 - oxy / (oxy + m.kinO2anox)) * (1.0 - no3 / (no3 + m.kinNO3anox));
    const double Rescale = 1.0 / (Oxicminlim + Denitrilim + Anoxiclim);

    const double CprodL = m.rLabile * ldetC;
    const double CprodS = m.rSemilabile * sdetC;
    double Cprod = CprodL + CprodS;
    Cprod = (Cprod > m.CprodMax) ? m.CprodMax : Cprod;
    const double Nprod = CprodL * m.NCrLdet + CprodS * m.NCrSdet;

    const double radsP = m.PAds_rS * po4 * fmax(odu, m.PAdsODU);
    const double rP = m.rLabile * (1.0 - Oxicminlim);
    const double Pprod = rP * detP;

    const double OxicMin = Cprod * Oxicminlim * Rescale;
    const double Denitrific = Cprod * Denitrilim * Rescale;
    const double AnoxicMin = Cprod * Anoxiclim * Rescale;

    const double Nitri = fT * m.rnit * nh3 * oxy / (oxy + m.ksO2nitri + relaxO2 * (ldetC + odu));
    const double OduOx = fT * m.rODUox * odu * oxy / (oxy + m.ksO2oduox + relaxO2 * (nh3 + ldetC));

    r[0] = -fT * CprodL;
    r[1] = -fT * CprodS;
    r[2] = fT * (radsP - Pprod);
    r[3] = fT * (Pprod - radsP);
    r[4] = -0.8 * Denitrific + Nitri;
    r[5] = (Nprod - Nitri) * m.rNH3Ads;
    r[6] = -OxicMin - 2.0 * Nitri - OduOx;
    r[7] = AnoxicMin - OduOx;
    if (denit) *denit = 0.8 * Denitrific;
}

// Zhang & Wirtz bioturbation factor of one cell, fabm_sediment_driver.F90:627-644
__device__ __forceinline__ double wtoc_cell(const KParams &p, double por, double poc0, double poc1)
{
    // weighted_toc + factor*porosity/(ones3d-porosity)*data, evaluated left to right (:627)
    double wt = 0.0;
    wt = wt + p.poc_factor[0] * por / (1.0 - por) * poc0;
    wt = wt + p.poc_factor[1] * por / (1.0 - por) * poc1;
    return wt;
}
__device__ __forceinline__ double bf3_cell(const KParams &p, int k, double wt, double avg)
{
    const double biomass = wt * p.e1[k] * avg / (p.L1 + p.L2 * p.e2[k]);
    return p.beta * pow(biomass, p.b) / wt;
}

template <int OP> struct OpTraits {
    static constexpr bool stepping = (OP != OP_RHS);
    static constexpr bool reads_base = (OP == OP_RK4_S2 || OP == OP_RK4_S3 || OP == OP_RK4_S4 ||
                                        OP == OP_RK38_S2 || OP == OP_RK38_S3 || OP == OP_RK38_S4);
    static constexpr bool final_stage = (OP == OP_EULER || OP == OP_ADAPTIVE || OP == OP_RK4_S4 ||
                                         OP == OP_RK38_S4);
    // stage evaluates the RHS on the scratch state c1 (buf[1-cur]) rather than on conc
    static constexpr bool in_is_c1 = reads_base;
};

// ---------------------------------------------------------------------------------------------
// the column kernel
// ---------------------------------------------------------------------------------------------
template <int MODEL, int OP, bool PROFILE3>
__global__ void __launch_bounds__(COL_BLOCK, COL_MIN_BLOCKS)
column_kernel(const __grid_constant__ KParams p)
{
    using T = OpTraits<OP>;
    int cur = 0;
    double dt = p.dt;
    bool final_sub = true, do_clip = false;
    if (T::stepping && p.use_ctl) {
        const Ctl *ctl = p.ctl;
        if (ctl->stop || ctl->steps_done >= ctl->steps_target) return;
        cur = ctl->cur;
        do_clip = ctl->do_clip != 0;
        if (OP == OP_ADAPTIVE) {
            dt = ctl->dt_red;
            final_sub = !(ctl->dt_int + dt < ctl->dt);
        } else {
            dt = ctl->dt;
        }
    }
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= p.ncol) return;

    const int K = p.K;
    const size_t ld = p.ld;
    const bool masked = p.mask[col] != 0;

    if (OP == OP_RHS) {
        if (masked) {  // driver :703-709 ; dissolved fluxes of masked columns are defined as 0
            for (int n = 0; n < NV; ++n)
                for (int k = 0; k < K; ++k) p.rhs_out[(size_t)(n * K + k) * ld + col] = 0.0;
            for (int n = NPART; n < NV; ++n) p.fluxes[(size_t)n * ld + col] = 0.0;
            return;
        }
    } else if (masked) {
        return;  // conc stays missing_value in both buffers; rhs == 0 there
    }

    const double *in = (T::in_is_c1 ? p.buf[1 - cur] : p.buf[cur]) + col;
    const double *base = p.buf[cur] + col;
    double *out = ((OP == OP_RK4_S4 || OP == OP_RK38_S4) ? p.buf[cur] : p.buf[1 - cur]) + col;
    const double *poc = p.buf[cur] + col;  // poc_classes%data => original conc (driver :476)
    double *aux1 = p.aux1 + col, *aux2 = p.aux2 + col;
    const double *por = p.por + col;

    if (MODEL == MSED_MODEL_TEST_SOLVER) {
        // rhs(i,j,k,:) = (i+j+k)*1.0d-8, src/test/test_Solver.F90:40
        const int i1 = col % p.inum + 1 + p.i_offset, j1 = col / p.inum + 1 + p.j_offset;
        for (int k = 0; k < K; ++k) {
            const double rhs = (double)(i1 + j1 + k + 1) * 1.0e-8;
            for (int n = 0; n < NV; ++n) {
                const size_t q = (size_t)(n * K + k) * ld;
                if (OP == OP_RHS) p.rhs_out[q + col] = rhs;
                else out[q] = __dadd_rn(in[q], __dmul_rn(dt, rhs));
            }
        }
        return;
    }

    // ---- per-column constants -----------------------------------------------------------
    const double temp = ld_ro(p.bdys + col);  // temp3d(:,:,k) = bdys(:,:,1), driver :602
    double cpart, fT = 1.0;
    if (PROFILE3) {
        cpart = 1.0 / 86400.0 / 10000.0;  // f_T = 1, bioturbation = 1, driver :622-623
    } else {
        const double f_T = exp(-4500.0 * (1.0 / (temp + 273.0) - (1.0 / 288.0)));  // :648
        cpart = p.bioturbation * f_T / 86400.0 / 10000.0;                          // :652
    }
    const double cdiss = (p.diffusivity + temp * 0.035) / 86400.0 / 10000.0;       // :682-683
    if (MODEL == MSED_MODEL_OMEXDIA_P)
        fT = exp(-p.om.E_a * (1.0 / (temp + 273.15) - 1.0 / 288.15));

    double avg_wt = 0.0;
    if (PROFILE3) {  // column integral of the weighted TOC, driver :629-633
        double s = 0.0;
        for (int k = 0; k < K; ++k) {
            const double pk = ld_ro(por + (size_t)k * ld);
            s += p.dz[k] * wtoc_cell(p, pk, poc[(size_t)k * ld], poc[(size_t)(K + k) * ld]);
        }
        avg_wt = s / p.cumdepth_last;
    }

    // ---- layer 1 and the upper boundary ---------------------------------------------------
    double cA[NV], cB[NV], F[NV];
    double porA, porB = 0.0;
#pragma unroll
    for (int n = 0; n < NV; ++n) cA[n] = ld_state(in + (size_t)(n * K) * ld);
    porA = ld_ro(por);

    double rest[NPART];
    bool casc[NPART];
    double cap_prev = 0.0;
    {
        double bf0 = p.bf[0];
        if (PROFILE3) bf0 = bf3_cell(p, 0, wtoc_cell(p, porA, poc[0], poc[(size_t)K * ld]), avg_wt);
        const double Dp = cpart * (1.0 - porA) * bf0;  // intf_porosity(:,:,1) = porosity(:,:,1), :434
        const double Dd = Dp + cdiss * porA;
        const double rdz0 = 1.0 / p.dz[0];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const bool part = n < NPART;
            const int bc = part ? p.bcup_part : p.bcup_diss;
            double f = 0.0;
            if (bc == 1 || bc == 4) {
                f = ld_ro(p.fluxes + (size_t)n * ld + col);  // driver :783,:792
            } else if (bc == 2) {                             // :786
                const double Cup = ld_ro(p.bdys + (size_t)(n + 1) * ld + col);
                const double C1 = part ? cA[n] * porA : cA[n];
                f = -(part ? Dp : Dd) * (C1 - Cup) * rdz0;
            }
            F[n] = f;
            if (!part) p.fluxes[(size_t)n * ld + col] = f;   // fluxes(:,:,n) = intFlux(:,:,1), :692
        }
        if (p.bcup_part == 4) {  // :792-794
            cap_prev = p.pom_flux_rate * (1.0 - porA) * p.dz[0];
#pragma unroll
            for (int n = 0; n < NPART; ++n) {
                rest[n] = F[n] - cap_prev;
                casc[n] = true;
                if (K == 1) F[n] += rest[n];  // k=2 > knum: :800-802
            }
        }
    }

    bool viol = false, nanf = false;

    // one layer: loads layer k+1 into (cn,porn), finishes layer k held in (cc,porc)
    auto layer = [&](const int k, double (&cc)[NV], double &porc, double (&cn)[NV], double &porn) {
        const bool has_next = (k + 1 < K);
        double basev[NV], a1[NV], a2[NV];
        if (has_next) {
#pragma unroll
            for (int n = 0; n < NV; ++n) cn[n] = ld_state(in + (size_t)(n * K + k + 1) * ld);
            porn = ld_ro(por + (size_t)(k + 1) * ld);
        }
        if (T::reads_base) {
#pragma unroll
            for (int n = 0; n < NV; ++n) basev[n] = ld_state(base + (size_t)(n * K + k) * ld);
        }
        if (OP == OP_RK4_S2 || OP == OP_RK4_S3 || OP == OP_RK4_S4 || OP == OP_RK38_S2 ||
            OP == OP_RK38_S3) {
#pragma unroll
            for (int n = 0; n < NV; ++n) a1[n] = aux1[(size_t)(n * K + k) * ld];
        }
        if (OP == OP_RK38_S3 || OP == OP_RK38_S4) {
#pragma unroll
            for (int n = 0; n < NV; ++n) a2[n] = aux2[(size_t)(n * K + k) * ld];
        }

        // local reaction rates (fabm_do, driver :700)
        double r[NV];
        if (MODEL == MSED_MODEL_OMEXDIA_P) {
            omexdia_rates(p.om, cc, fT, r, nullptr);
        } else {
#pragma unroll
            for (int n = 0; n < NV; ++n) r[n] = 0.0;
        }

        // flux through the lower interface (diff3d :776-778; BcDown = 3, :590,:813)
        double Fn[NV];
        if (has_next) {
            const double intf = 0.5 * (porc + porn);  // :435
            double bfk = p.bf[k + 1];
            if (PROFILE3)
                bfk = bf3_cell(p, k + 1,
                               wtoc_cell(p, porn, poc[(size_t)(k + 1) * ld], poc[(size_t)(K + k + 1) * ld]),
                               avg_wt);
            const double Dp = cpart * (1.0 - intf) * bfk;
            const double Dd = Dp + cdiss * intf;
            const double rdzc = p.rdzc[k];
            const double mDp = -Dp * rdzc, mDd = -Dd * rdzc;
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                if (n < NPART) Fn[n] = mDp * (cn[n] * porn - cc[n] * porc);
                else Fn[n] = mDd * (cn[n] - cc[n]);
            }
            if (p.bcup_part == 4) {  // distributed POM flux cascade, :795-802 (kk = k+2, 1-based)
                double cap = p.pom_flux_rate * (1.0 - porn) * p.dz[k + 1];
                if (k + 2 > 2 && cap > cap_prev) cap = cap_prev;  // driver :285-291
                cap_prev = cap;
#pragma unroll
                for (int n = 0; n < NPART; ++n) {
                    if (casc[n] && rest[n] > 0.0) {
                        Fn[n] += rest[n];
                        rest[n] -= cap;
                        if (k + 2 == K) Fn[n] += rest[n];
                    } else {
                        casc[n] = false;
                    }
                }
            }
        } else {
#pragma unroll
            for (int n = 0; n < NV; ++n) Fn[n] = 0.0;
        }

        // dC (:819) with the particulate rescaling (:677-678) folded: both reduce to
        // (Flux(k)-Flux(k+1)) / (porosity*dz)
        const double rpd = 1.0 / (porc * p.dz[k]);
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const double rhs = r[n] + (F[n] - Fn[n]) * rpd;  // driver :715
            F[n] = Fn[n];
            const size_t q = (size_t)(n * K + k) * ld;
            const double c0 = cc[n];
            double newc = 0.0;
            if (OP == OP_RHS) {
                p.rhs_out[q + col] = rhs;
            } else if (OP == OP_EULER || OP == OP_ADAPTIVE) {
                newc = __dadd_rn(c0, __dmul_rn(dt, rhs));             // :102,:111
                if (OP == OP_ADAPTIVE)                                  // :121
                    viol |= (__dsub_rn(newc, __dmul_rn(p.fac, c0)) < 0.0);
            } else if (OP == OP_RK4_S1) {                               // :147
                newc = __dadd_rn(c0, __dmul_rn(0.5 * dt, rhs));
                aux1[q] = 0.5 * rhs;
            } else if (OP == OP_RK4_S2) {                               // :152
                newc = __dadd_rn(basev[n], __dmul_rn(0.5 * dt, rhs));
                aux1[q] = __dadd_rn(a1[n], rhs);
            } else if (OP == OP_RK4_S3) {                               // :156
                newc = __dadd_rn(basev[n], __dmul_rn(dt, rhs));
                aux1[q] = __dadd_rn(a1[n], rhs);
            } else if (OP == OP_RK4_S4) {                               // :160
                const double third = 1.0 / 3.0;
                newc = __dadd_rn(basev[n], __dmul_rn(dt * third, __dadd_rn(a1[n], 0.5 * rhs)));
            } else if (OP == OP_RK38_S1) {                              // :169
                const double third = 1.0 / 3.0;
                newc = __dadd_rn(c0, __dmul_rn(third * dt, rhs));
                aux1[q] = rhs;
            } else if (OP == OP_RK38_S2) {                              // :174
                const double third = 1.0 / 3.0;
                const double r0 = a1[n];
                newc = __dadd_rn(basev[n], __dmul_rn(dt, __dsub_rn(rhs, __dmul_rn(third, r0))));
                aux1[q] = __dsub_rn(r0, rhs);
                aux2[q] = __dadd_rn(r0, __dmul_rn(3.0, rhs));
            } else if (OP == OP_RK38_S3) {                              // :178
                newc = __dadd_rn(basev[n], __dmul_rn(dt, __dadd_rn(a1[n], rhs)));
                aux2[q] = __dadd_rn(a2[n], __dmul_rn(3.0, rhs));
            } else if (OP == OP_RK38_S4) {                              // :182
                newc = __dadd_rn(basev[n], __dmul_rn(dt * 1.0 / 8.0, __dadd_rn(a2[n], rhs)));
            }
            if (OP != OP_RHS) {
                if (T::final_stage && do_clip && final_sub) {
                    nanf |= (newc != newc);                             // component :2392
                    newc = (newc < p.om.minimum[n]) ? p.om.minimum[n] : newc;  // :1728-1730
                }
                out[q] = newc;
            }
        }
    };

    for (int k = 0; k < K; k += 2) {
        layer(k, cA, porA, cB, porB);
        if (k + 1 < K) layer(k + 1, cB, porB, cA, porA);
    }

    if (OP == OP_ADAPTIVE && viol) atomicOr(&p.ctl->flags[0], 1);
    if (T::final_stage && nanf) atomicOr(&p.ctl->flags[1], 1);
}

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
        const double v = p0 * (1.0 - porosity_fac * (zc[k] - zc[0]));
        por[(size_t)k * ld + col] = m ? 1.0 : v;
    }
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
