// msed_chain.cuh -- included inside namespace msed after msed_pair.cuh.
//
// A chain of consecutive Euler / adaptive-Euler steps with the column state held in REGISTERS: one
// warp owns one sediment column, lane k owns layer k (knum <= 32) -- or, LPL = 2, layers 2k and 2k+1
// (32 < knum <= 64: BASELINE config 4's 40 layers on a tile too small for a thread per column).  The
// vertical stencil of diff3d is one layer wide, so a step needs the state of the layer below the lane's
// deepest layer (one __shfl_down per variable) and the flux through the upper interface of its first
// layer, which the lane above has just computed (one __shfl_up per variable); the interface between a
// lane's own two layers is local.  The state is read from HBM once, advanced nsub steps, and written once: per step the
// chain moves 128/nsub bytes per cell instead of the 128 of the single-step kernel.
//
// Why a second fused kernel next to pair_kernel (thread per column, two steps per launch):
//  - small tiles (BASELINE config 2: 100x100 columns) give the thread-per-column kernels 79 CTAs, less
//    than one warp per SM scheduler, and every thread walks 30 layers x 282 dependent-ish instructions
//    per step; with a warp per column the same tile is 10,000 warps, enough to fill the machine, and a
//    whole coupling interval (10 steps) is one launch;
//  - land columns cost nothing (a masked warp exits; thread-per-column warps idle their land lanes);
//  - per-layer coefficients (porosity, interface diffusivities, 1/(porosity*dz)) are step-invariant
//    and stay in registers for the whole chain.
//
// Semantics are those of nsub ode_solver calls (solver_library.F90:104-140), each followed by the
// component's check_NaN / minimum clip (fabm_sediment_component.F90:1718-1732): a step's output is
// clipped before the next step reads it, every step raises the violation / NaN flags, and
// plan_controller_kernel commits the chain only if the reference would have taken exactly the planned
// decisions.  The plan (msed.cu run_steps) may hold sub-cycled steps (SUB): every call then runs at
// dt_acc = dt/4^depth, i.e. its first sub-step also evaluates -- with the same RHS, the state being unchanged
// by a rejection -- the attempts at dt, dt/4, .. that the reference rejects (:126-128) and raises "rejection
// seen" for each, 4^depth sub-steps follow each other without a clip, and the clip comes after the last.
// Otherwise nothing is committed (the input buffer is untouched) and the host redoes the same attempts
// with the single-step kernel -- speculation with a free rollback, exactly as for pairs.  The
// arithmetic is the shared inline code of msed_column.cuh, so a committed chain is bit-identical to
// the single attempts.
//
// Same scope as pair_kernel: bcup_particulate = 1, bioturbation_profile != 3, closed-form porosity.

constexpr int CHAIN_WARPS = 8;                  // columns per CTA: 8 adjacent columns = 64 contiguous bytes per row
constexpr int CHAIN_BLOCK = CHAIN_WARPS * 32;
constexpr int CHAIN_MAX_LAYERS = 64;            // 32 layers per LPL
#ifndef MSED_CHAIN_MIN_BLOCKS
#define MSED_CHAIN_MIN_BLOCKS 2
#endif
#ifndef MSED_CHAIN_MAX_STEPS
#define MSED_CHAIN_MAX_STEPS 16                 // steps per launch: bounds the work a failed speculation throws away
#endif

// CLIP: the component wrapper (check_NaN + minimum clip after every step) is on -- a host-side fact
// (msed_step / msed_run set it, msed_ode_solver does not), mirrored in Ctl::do_clip
// SUB: the plan holds sub-cycled steps (KParams::depth > 0); a separate instantiation so that the common
// chain (every step accepted at dt) keeps its straight step loop
// LPL: layers per lane (1: knum <= 32, 128 registers, two CTAs per SM; 2: knum <= 64, one CTA per SM)
template <int MODEL, bool ADAPTIVE, bool CLIP, bool SUB, int LPL = 1>
__global__ void __launch_bounds__(CHAIN_BLOCK, LPL == 1 ? MSED_CHAIN_MIN_BLOCKS : 1)
chain_kernel(const __grid_constant__ KParams p, const int nsub)
{
    const Ctl *ctl = p.ctl;
    // the plan was made for one definite control state (see pair_kernel)
    if (ctl->stop || ctl->pairs_disabled || ctl->steps_done != p.gate_steps) return;
    const int cur = ctl->cur;
    const double dt = p.dt_acc;
    const int depth = SUB ? p.depth : 0;
    const int nq = 1 << (2 * depth);               // accepted sub-steps per ode_solver call
    const double dt_up[MAX_PLAN_DEPTH] = {dt * 4.0, dt * 16.0};   // the rejected step sizes, exact (:127)
    unsigned upmask = 0;                           // bit s*depth + l: step s saw the rejection at dt_up[l]
    // a violation can only be rejected while dt_red > dt_min (solver_library.F90:126); a chain whose
    // violation flag is already up cannot be committed, so warps that start later skip their work
    const bool rejectable = ADAPTIVE && dt > ctl->dt_min;
    const volatile int *rflags = ctl->flags;
    if (rejectable && (rflags[0] | rflags[2])) return;

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int col = p.col0 + blockIdx.x * CHAIN_WARPS + (threadIdx.x >> 5);
    if (col >= p.col_end) return;  // warp-uniform from here on: the shuffles below see all 32 lanes
    if (p.mask[col] != 0) return;  // conc stays missing_value in both buffers

    const int K = p.K;
    const size_t ld = p.ld;
    const size_t plane = (size_t)K * ld;
    // the lane's layers, top to bottom; spare slots shadow the deepest layer, they never store or flag
    bool active[LPL], has_next[LPL];
    int k[LPL];
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
        const int kj = LPL * lane + j;
        active[j] = kj < K;
        k[j] = active[j] ? kj : K - 1;
        has_next[j] = kj + 1 < K;
    }

    double cc[LPL][NV];
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
        const double *in = p.buf[cur] + (size_t)k[j] * ld + col;
#pragma unroll
        for (int n = 0; n < NV; ++n) cc[j][n] = in[(size_t)n * plane];
    }

    // ---- step-invariant coefficients of the lane's layers and of their lower interfaces ------------
    const double por_surf = (p.por_mode == 2) ? ld_ro(p.por + col) : 1.0;
    const double temp = ld_ro(p.bdys + col);
    double cpart, cdiss, fT;
    column_constants<MODEL, false>(p, temp, cpart, cdiss, fT);
    double fT_diag = fT;  // what field_kernel reports for the reaction-free model
    if (MODEL != MSED_MODEL_OMEXDIA_P && p.denit_out)
        fT_diag = exp(-p.om.E_a * (1.0 / (temp + 273.15) - 1.0 / 288.15));

    double porc[LPL], porn[LPL], mDp[LPL], mDd[LPL], rpd[LPL];
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
        porc[j] = __dmul_rn(por_surf, p.portab[k[j]]);
        porn[j] = mDp[j] = mDd[j] = 0.0;
        if (has_next[j]) {
            porn[j] = __dmul_rn(por_surf, p.portab[k[j] + 1]);
            interface_coeffs(cpart, cdiss, porc[j], porn[j], p.bf[k[j] + 1], p.rdzc[k[j]], mDp[j], mDd[j]);
        }
        rpd[j] = fast_rcp(MSED_MUL(porc[j], p.dz[k[j]]));
    }
    // upper boundary (used by lane 0 only): diff3d :782-803
    const int bc_diss = p.bcup_diss;
    const double por0 = __dmul_rn(por_surf, p.portab[0]);
    double Dp0, Dd0;
    top_coeffs(cpart, cdiss, por0, p.bf[0], Dp0, Dd0);
    const double rdz0 = 1.0 / p.dz[0];

    int viol[LPL];  // sign bit = some relative change fell below relative_change_min (violates_acc)
    bool nanf[LPL];
    double dn_last[LPL];
#pragma unroll
    for (int j = 0; j < LPL; ++j) { viol[j] = 0; nanf[j] = false; dn_last[j] = 0.0; }

    // where lane 0 finds the upper-boundary input of a dissolved variable: the concentration above the
    // bed for BcUp = 2 (bdys row n+1), the imposed flux for BcUp = 1 (fluxes row n); any other BcUp
    // reads nothing, the address only has to be valid
    const double *top_part = p.fluxes + col;
    const double *top_diss = (bc_diss == 2) ? p.bdys + ld + col : p.fluxes + col;

    const int nacc = nsub * nq;
#ifdef MSED_CHAIN_UNROLL
    MSED_UNROLL_PRAGMA(MSED_CHAIN_UNROLL)
#endif
    for (int a = 0; a < nacc; ++a) {
        const int s = SUB ? a / nq : a, q = SUB ? a - s * nq : 0;
        const bool last = (a == nacc - 1);
        const bool final_sub = !SUB || q == nq - 1;   // this sub-step ends an ode_solver call

        // upper-boundary inputs, read by every lane (one address per warp: a broadcast) and not inside a
        // lane-0 branch: they are step-invariant, so the compiler lifts them out of the loop and keeps them in
        // registers (inside the branch they were an L1 round trip per step on the critical path of lane 0)
        double tin[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) tin[n] = ld_ro((n < NPART ? top_part : top_diss) + (size_t)n * ld);

        // the state of the layer below the lane's deepest layer (the next lane's first) is requested first: the
        // shuffles complete while the reaction term, which does not need it, is evaluated
        double cn[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) cn[n] = __shfl_down_sync(FULL, cc[0][n], 1);

        // local reaction rates (fabm_do, driver :700): independent of the neighbours
        double r[LPL][NV];
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
            double dn = 0.0;
            if (MODEL == MSED_MODEL_OMEXDIA_P) {
                omexdia_rates(p.om, cc[j], fT, r[j], &dn);
            } else {
                if (last && p.denit_out) omexdia_rates(p.om, cc[j], fT_diag, r[j], &dn);
#pragma unroll
                for (int n = 0; n < NV; ++n) r[j][n] = 0.0;
            }
            if (last) dn_last[j] = dn;  // the FABM diagnostic describes the state of the last get_rhs call
        }

        // flux through the lower interface of every layer of the lane (diff3d :776-778; BcDown = 3 below the
        // deepest layer): towards the lane's own next layer, or towards the next lane's first
        double Fn[LPL][NV];
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const double below = (j + 1 < LPL) ? cc[j + 1 < LPL ? j + 1 : j][n] : cn[n];
                const double f = (n < NPART) ? flux_particulate(mDp[j], below, porn[j], cc[j][n], porc[j])
                                             : flux_dissolved(mDd[j], below, cc[j][n]);
                Fn[j][n] = has_next[j] ? f : 0.0;
            }
        }
        // flux through the upper interface of the lane's first layer = the lower-interface flux of the lane
        // above's last layer.  (Shuffling the state up instead and recomputing that flux here, to drop this
        // second exchange from the critical path, was measured 7-8 % slower: 19 more fp64 instructions per step.)
        double F[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) F[n] = __shfl_up_sync(FULL, Fn[LPL - 1][n], 1);
        // upper boundary, diff3d :782-803: only lane 0 keeps the result.  Every lane evaluates it (a warp
        // pays for a one-lane branch body anyway; selects keep the instruction stream straight).
        // Particulates: BcUp = 1 (the host only fuses configurations without the distributed POM flux
        // cascade), Flux(1) = the imposed sinking flux (:783)
        const bool top = (lane == 0);
#pragma unroll
        for (int n = 0; n < NPART; ++n) F[n] = top ? tin[n] : F[n];
        if (bc_diss == 2) {                                     // Dirichlet, :786
#pragma unroll
            for (int n = NPART; n < NV; ++n) {
                const double f = top_flux_dirichlet(Dd0, cc[0][n], tin[n], rdz0);
                F[n] = top ? f : F[n];
            }
        } else if (bc_diss == 1 || bc_diss == 4) {              // imposed flux (rewritten unchanged below)
#pragma unroll
            for (int n = NPART; n < NV; ++n) F[n] = top ? tin[n] : F[n];
        } else {
            // BcUp = 3: explicit zero (:789).  Any other value: diff3d never assigns Flux(1), which
            // keeps what the previous variable left in get_rhs's intFlux -- see column_kernel
            const double f = (bc_diss == 3) ? 0.0 : tin[NPART - 1];
#pragma unroll
            for (int n = NPART; n < NV; ++n) F[n] = top ? f : F[n];
        }
        if (last && top) {
#pragma unroll
            for (int n = NPART; n < NV; ++n) p.fluxes[(size_t)n * ld + col] = F[n];  // :692
        }

        int vup = 0;
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
            double raw[NV];  // new state before the clip (check_NaN looks at it first, component :1718)
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const double fup = (j == 0) ? F[n] : Fn[j > 0 ? j - 1 : 0][n];
                const double rhs = layer_rhs(fup, Fn[j][n], rpd[j], r[j][n]);
                const double c0 = cc[j][n];
                double newc = euler_update(dt, rhs, c0);
                if (ADAPTIVE) violates_acc(viol[j], p.fac, c0, newc);
                if (ADAPTIVE && SUB && q == 0) {   // the planned rejections: tested exactly as a single attempt does
#pragma unroll
                    for (int l = 0; l < MAX_PLAN_DEPTH; ++l)
                        if (l < depth && active[j] && violates(p.fac, c0, euler_update(dt_up[l], rhs, c0))) vup |= 1 << l;
                }
                raw[n] = newc;
                if (CLIP && final_sub) {
                    if (n & 1) nanf[j] |= either_nan(raw[n - 1], raw[n]);
                    const double mn = p.om.minimum[n];
                    newc = clip_min(newc, mn);
                }
                cc[j][n] = newc;
            }
        }
        if (ADAPTIVE && SUB && q == 0) {
#pragma unroll
            for (int l = 0; l < MAX_PLAN_DEPTH; ++l)
                if (l < depth && __any_sync(FULL, (vup & (1 << l)) != 0)) upmask |= 1u << (s * depth + l);
        }
        // a rejectable violation anywhere in the column: the chain will not be committed, stop here
        bool v_any = false;
#pragma unroll
        for (int j = 0; j < LPL; ++j) v_any |= (viol[j] < 0 && active[j]);
        if (rejectable && __any_sync(FULL, v_any)) {
            if (lane == 0) {
                atomicOr(&p.ctl->flags[0], 1);
                atomicMax(&p.ctl->flags[FLAG_FAIL], 64 - s);   // where: the earliest such step wins (run_steps re-plans)
            }
            return;
        }
    }

    bool v_any = false, n_any = false;
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
        if (active[j]) {
            double *out = p.buf[1 - cur] + (size_t)k[j] * ld + col;
#pragma unroll
            for (int n = 0; n < NV; ++n) out[(size_t)n * plane] = cc[j][n];
            if (p.denit_out) p.denit_out[(size_t)k[j] * ld + col] = dn_last[j];
        }
        v_any |= (viol[j] < 0 && active[j]);
        n_any |= (nanf[j] && active[j]);
    }
    const bool any_viol = __any_sync(FULL, v_any);
    const bool any_nan = __any_sync(FULL, n_any);
    if (lane == 0) {
        if (ADAPTIVE && any_viol) atomicOr(&p.ctl->flags[0], 1);
        if (any_nan) atomicOr(&p.ctl->flags[1], 1);
        for (int b = 0; SUB && upmask; ++b, upmask >>= 1)
            if (upmask & 1u) atomicOr(&p.ctl->flags[FLAG_UP0 + p.up_slot + b], 1);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same for the Runge-Kutta integrators (solver_library.F90:142-185): nsub ode_solver calls per launch, the four
// stages of every call evaluated on the column in registers -- the stage state, the base state and the accumulators
// never leave the warp, where the thread-per-column path (rk_pair_kernel) moves them through HBM between its two
// launches per call and, on a tile too small to fill the machine, walks every column serially.  Stage formulas and
// their rounding are those of column_kernel's OP_RK4_* / OP_RK38_* stages, so a committed chain is bit-identical to
// the staged path.  There is no accept decision: the launch is committed unless check_NaN (component :1718) fires.
template <int MODEL, int METHOD, bool CLIP, int LPL = 1>
__global__ void __launch_bounds__(CHAIN_BLOCK, 1)
rk_chain_kernel(const __grid_constant__ KParams p, const int nsub)
{
    const Ctl *ctl = p.ctl;
    if (ctl->stop || ctl->pairs_disabled || ctl->steps_done != p.gate_steps) return;
    const int cur = ctl->cur;
    const double dt = p.dt_acc;

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int col = p.col0 + blockIdx.x * CHAIN_WARPS + (threadIdx.x >> 5);
    if (col >= p.col_end) return;  // warp-uniform from here on
    if (p.mask[col] != 0) return;

    const int K = p.K;
    const bool top = lane == 0;
    const size_t ld = p.ld;
    const size_t plane = (size_t)K * ld;
    bool active[LPL], has_next[LPL];
    int k[LPL];
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
        const int kj = LPL * lane + j;
        active[j] = kj < K;
        k[j] = active[j] ? kj : K - 1;
        has_next[j] = kj + 1 < K;
    }
    double cc[LPL][NV];
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
        const double *in = p.buf[cur] + (size_t)k[j] * ld + col;
#pragma unroll
        for (int n = 0; n < NV; ++n) cc[j][n] = in[(size_t)n * plane];
    }
    const double por_surf = (p.por_mode == 2) ? ld_ro(p.por + col) : 1.0;
    const double temp = ld_ro(p.bdys + col);
    double cpart, cdiss, fT;
    column_constants<MODEL, false>(p, temp, cpart, cdiss, fT);
    double fT_diag = fT;
    if (MODEL != MSED_MODEL_OMEXDIA_P && p.denit_out)
        fT_diag = exp(-p.om.E_a * (1.0 / (temp + 273.15) - 1.0 / 288.15));
    double porc[LPL], porn[LPL], mDp[LPL], mDd[LPL], rpd[LPL];
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
        porc[j] = __dmul_rn(por_surf, p.portab[k[j]]);
        porn[j] = mDp[j] = mDd[j] = 0.0;
        if (has_next[j]) {
            porn[j] = __dmul_rn(por_surf, p.portab[k[j] + 1]);
            interface_coeffs(cpart, cdiss, porc[j], porn[j], p.bf[k[j] + 1], p.rdzc[k[j]], mDp[j], mDd[j]);
        }
        rpd[j] = fast_rcp(MSED_MUL(porc[j], p.dz[k[j]]));
    }
    const int bc_diss = p.bcup_diss;
    const double por0 = __dmul_rn(por_surf, p.portab[0]);
    double Dp0, Dd0;
    top_coeffs(cpart, cdiss, por0, p.bf[0], Dp0, Dd0);
    const double rdz0 = 1.0 / p.dz[0];
    const double *top_part = p.fluxes + col;
    const double *top_diss = (bc_diss == 2) ? p.bdys + ld + col : p.fluxes + col;
    double tin[NV];
#pragma unroll
    for (int n = 0; n < NV; ++n) tin[n] = ld_ro((n < NPART ? top_part : top_diss) + (size_t)n * ld);

    double Ftop[NV];       // Flux(1) of the last RHS evaluation (lane 0): sed%fluxes(dissolved), driver :692
    double dn_last[LPL];   // FABM denit diagnostic of the last RHS evaluation
#pragma unroll
    for (int j = 0; j < LPL; ++j) dn_last[j] = 0.0;
    // get_rhs for the state x of this lane's layers
    auto rhs_of = [&](const double (&x)[LPL][NV], double (&rhs)[LPL][NV], bool want_dn) __attribute__((always_inline)) {
        double cn[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) cn[n] = __shfl_down_sync(FULL, x[0][n], 1);
        double r[LPL][NV], Fn[LPL][NV], F[NV];
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
            double dn = 0.0;
            if (MODEL == MSED_MODEL_OMEXDIA_P) {
                omexdia_rates(p.om, x[j], fT, r[j], &dn);
            } else {
                if (want_dn && p.denit_out) omexdia_rates(p.om, x[j], fT_diag, r[j], &dn);
#pragma unroll
                for (int n = 0; n < NV; ++n) r[j][n] = 0.0;
            }
            if (want_dn) dn_last[j] = dn;
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const double below = (j + 1 < LPL) ? x[j + 1 < LPL ? j + 1 : j][n] : cn[n];
                const double f = (n < NPART) ? flux_particulate(mDp[j], below, porn[j], x[j][n], porc[j])
                                             : flux_dissolved(mDd[j], below, x[j][n]);
                Fn[j][n] = has_next[j] ? f : 0.0;
            }
        }
#pragma unroll
        for (int n = 0; n < NV; ++n) F[n] = __shfl_up_sync(FULL, Fn[LPL - 1][n], 1);
#pragma unroll
        for (int n = 0; n < NPART; ++n) F[n] = top ? tin[n] : F[n];
        if (bc_diss == 2) {
#pragma unroll
            for (int n = NPART; n < NV; ++n) {
                const double f = top_flux_dirichlet(Dd0, x[0][n], tin[n], rdz0);
                F[n] = top ? f : F[n];
            }
        } else if (bc_diss == 1 || bc_diss == 4) {
#pragma unroll
            for (int n = NPART; n < NV; ++n) F[n] = top ? tin[n] : F[n];
        } else {
            const double f = (bc_diss == 3) ? 0.0 : tin[NPART - 1];
#pragma unroll
            for (int n = NPART; n < NV; ++n) F[n] = top ? f : F[n];
        }
#pragma unroll
        for (int n = 0; n < NV; ++n) Ftop[n] = F[n];
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const double fup = (j == 0) ? F[n] : Fn[j > 0 ? j - 1 : 0][n];
                rhs[j][n] = layer_rhs(fup, Fn[j][n], rpd[j], r[j][n]);
            }
        }
    };

    bool nanf = false;
    const double third = 1.0 / 3.0;
#define MSED_RKC_ALL(stmt) _Pragma("unroll") for (int j = 0; j < LPL; ++j) _Pragma("unroll") for (int n = 0; n < NV; ++n) { stmt; }
    for (int s = 0; s < nsub; ++s) {
        const bool last = (s == nsub - 1);
        double rhs[LPL][NV], base[LPL][NV], c1[LPL][NV], a1[LPL][NV];
        MSED_RKC_ALL(base[j][n] = cc[j][n])
        if (METHOD == MSED_RUNGE_KUTTA_4) {                              // column_kernel OP_RK4_S1..S4
            rhs_of(cc, rhs, false);
            MSED_RKC_ALL(c1[j][n] = fma(0.5 * dt, rhs[j][n], base[j][n]); a1[j][n] = 0.5 * rhs[j][n])
            rhs_of(c1, rhs, false);
            MSED_RKC_ALL(c1[j][n] = fma(0.5 * dt, rhs[j][n], base[j][n]); a1[j][n] = a1[j][n] + rhs[j][n])
            rhs_of(c1, rhs, false);
            MSED_RKC_ALL(c1[j][n] = fma(dt, rhs[j][n], base[j][n]); a1[j][n] = a1[j][n] + rhs[j][n])
            rhs_of(c1, rhs, last);
            MSED_RKC_ALL(cc[j][n] = fma(dt * third, fma(0.5, rhs[j][n], a1[j][n]), base[j][n]))
        } else {                                                         // OP_RK38_S1..S4
            double a2[LPL][NV];
            rhs_of(cc, rhs, false);
            MSED_RKC_ALL(c1[j][n] = fma(third * dt, rhs[j][n], base[j][n]); a1[j][n] = rhs[j][n])
            rhs_of(c1, rhs, false);
            MSED_RKC_ALL(const double r0 = a1[j][n]; c1[j][n] = fma(dt, fma(-third, r0, rhs[j][n]), base[j][n]);
                         a1[j][n] = r0 - rhs[j][n]; a2[j][n] = fma(3.0, rhs[j][n], r0))
            rhs_of(c1, rhs, false);
            MSED_RKC_ALL(c1[j][n] = fma(dt, a1[j][n] + rhs[j][n], base[j][n]); a2[j][n] = fma(3.0, rhs[j][n], a2[j][n]))
            rhs_of(c1, rhs, last);
            MSED_RKC_ALL(cc[j][n] = fma(dt * 1.0 / 8.0, a2[j][n] + rhs[j][n], base[j][n]))
        }
        if (CLIP) {   // check_NaN on the new state, then the minimum clip (component :1718-1732)
#pragma unroll
            for (int j = 0; j < LPL; ++j)
#pragma unroll
                for (int n = 0; n < NV; ++n) {
                    if (n & 1) nanf |= active[j] && either_nan(cc[j][n - 1], cc[j][n]);
                }
            MSED_RKC_ALL(cc[j][n] = clip_min(cc[j][n], p.om.minimum[n]))
        }
    }
#undef MSED_RKC_ALL

#pragma unroll
    for (int j = 0; j < LPL; ++j)
        if (active[j]) {
            double *out = p.buf[1 - cur] + (size_t)k[j] * ld + col;
#pragma unroll
            for (int n = 0; n < NV; ++n) out[(size_t)n * plane] = cc[j][n];
            if (p.denit_out) p.denit_out[(size_t)k[j] * ld + col] = dn_last[j];
        }
    if (top && nsub > 0) {
#pragma unroll
        for (int n = NPART; n < NV; ++n) p.fluxes[(size_t)n * ld + col] = Ftop[n];   // :692
    }
    const bool any_nan = __any_sync(FULL, nanf);
    if (top && any_nan) atomicOr(&p.ctl->flags[1], 1);
}
