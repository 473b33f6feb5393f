// msed_spinup.cuh -- included inside namespace msed after msed_column.cuh.
//
// The 1-D pre-simulation of the component (fabm_sediment_component.F90:557-632) for a whole BATCH of
// independent columns in one launch: one warp per member, lane k = layer k (knum <= 32; LPL = 2: layers 2k and
// 2k+1, knum <= 64, as in chain_kernel), the column in registers for all nsteps ode_solver calls (dt_spinup =
// 3600 s, :574).  A member is what the reference calls
// sed1d: a 1x1xknum clone of the sediment with Dirichlet upper boundaries for the dissolved variables (:608),
// imposed particulate fluxes, constant bioturbation (:611) and adaptive_solver_diagnostics on (:610).  Members
// may differ in boundary values, reaction parameters and initial values -- an ensemble of parameter sets, or the
// spin-up of many forcing classes at start-up -- and each of them is its OWN domain: the accept test of
// adaptive Euler (solver_library.F90:121) is taken per member, so the sub-cycling of a stiff member costs the
// others nothing, and last_min_dt / last_min_dt_grid_cell (:130-135) are kept per member.  ode_solver is called
// bare, without the component's check_NaN / clip (:614-618).
//
// All four integrators of solver_library.F90:80-189; the arithmetic is the shared inline code of
// msed_column.cuh in the order chain_kernel uses it, so every member is bit-identical to
// msed_spinup_column (tests/test_gpu_spinup.py).

constexpr int SPINUP_WARPS = 4;                 // members per CTA
constexpr int SPINUP_BLOCK = SPINUP_WARPS * 32;


template <int MODEL, int LPL>
__global__ void __launch_bounds__(SPINUP_BLOCK)
spinup_kernel(const __grid_constant__ KParams p, const SpinupArgs a)
{
    __shared__ OmexDev som[SPINUP_WARPS];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int col = blockIdx.x * SPINUP_WARPS + w;       // member
    const int colc = col < p.ncol ? col : p.ncol - 1;    // spare warps shadow the last member (no exits before
                                                         // the barrier below)
    {
        const double *src = reinterpret_cast<const double *>(a.om ? a.om + colc : &p.om);
        double *dst = reinterpret_cast<double *>(&som[w]);
        for (int q = lane; q < (int)(sizeof(OmexDev) / sizeof(double)); q += 32) dst[q] = src[q];
    }
    __syncthreads();
    if (col >= p.ncol) return;
    const OmexDev &om = som[w];

    const int K = p.K;
    const bool top = lane == 0;
    const size_t ld = p.ld;
    const size_t plane = (size_t)K * ld;
    // the lane's layers, top to bottom; spare slots shadow the deepest layer, they never store
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
        const double *in = p.buf[0] + (size_t)k[j] * ld + col;
#pragma unroll
        for (int n = 0; n < NV; ++n) cc[j][n] = in[(size_t)n * plane];
    }
    // step-invariant coefficients (as chain_kernel)
    const double por_surf = (p.por_mode == 2) ? ld_ro(p.por + col) : 1.0;
    const double temp = ld_ro(p.bdys + col);
    double cpart, cdiss, fT;
    column_constants<MODEL, false>(p, temp, cpart, cdiss, fT);
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

    double Ftop[NV];   // Flux(1) of the last RHS evaluation (lane 0): sed%fluxes(dissolved), driver :692
    // get_rhs for the state x of this lane's layers
    auto rhs_of = [&](const double (&x)[LPL][NV], double (&rhs)[LPL][NV]) {
        double cn[NV];   // the state below the lane's deepest layer: the next lane's first layer
#pragma unroll
        for (int n = 0; n < NV; ++n) cn[n] = __shfl_down_sync(FULL, x[0][n], 1);
        double r[LPL][NV], Fn[LPL][NV], F[NV];
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
            if (MODEL == MSED_MODEL_OMEXDIA_P) {
                omexdia_rates(om, x[j], fT, r[j], nullptr);
            } else {
#pragma unroll
                for (int n = 0; n < NV; ++n) r[j][n] = 0.0;
            }
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

    double last_min_dt = a.last_min_dt0;
    int cell_k = -99, cell_n = -99;
    long long rhs_evals = 0, subcycles = 0;
    const double dt = a.dt, third = 1.0 / 3.0;

    for (long long s = 0; s < a.nsteps; ++s) {
        double rhs[LPL][NV];
        if (a.method == MSED_EULER) {                                   // solver_library.F90:99-102
            rhs_of(cc, rhs);
#pragma unroll
            for (int j = 0; j < LPL; ++j)
#pragma unroll
                for (int n = 0; n < NV; ++n) cc[j][n] = euler_update(dt, rhs[j][n], cc[j][n]);
            rhs_evals += 1;
        } else if (a.method == MSED_ADAPTIVE_EULER) {                   // :104-140
            double dt_int = 0.0, dt_red = dt;
            while (dt_int < dt) {
                rhs_of(cc, rhs);
                rhs_evals += 1;
                double c1[LPL][NV];
                bool viol = false;
#pragma unroll
                for (int j = 0; j < LPL; ++j)
#pragma unroll
                    for (int n = 0; n < NV; ++n) {
                        c1[j][n] = euler_update(dt_red, rhs[j][n], cc[j][n]);
                        viol |= active[j] && violates(p.fac, cc[j][n], c1[j][n]);   // :121
                    }
                if (__any_sync(FULL, viol) && dt_red > a.dt_min) {      // :126-128
                    dt_red = dt_red * 0.25;
                    subcycles += 1;
                    continue;
                }
                if (dt_red < last_min_dt) {                             // :130-135
                    last_min_dt = dt_red;
                    // minloc((c1-c)/c) in Fortran array order (k fastest, then n), NaNs skipped, first minimum
                    double bv = 0.0;
                    int bi = 0x7fffffff;
#pragma unroll
                    for (int n = 0; n < NV; ++n)
#pragma unroll
                        for (int j = 0; j < LPL; ++j) {
                            const double v = __ddiv_rn(__dsub_rn(c1[j][n], cc[j][n]), cc[j][n]);
                            const int idx = k[j] + K * n;
                            if (active[j] && v == v && (bi == 0x7fffffff || v < bv)) { bv = v; bi = idx; }
                        }
                    for (int o = 16; o > 0; o >>= 1) {
                        const double ov = __shfl_xor_sync(FULL, bv, o);
                        const int oi = __shfl_xor_sync(FULL, bi, o);
                        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
                    }
                    if (bi != 0x7fffffff) { cell_k = bi % K + 1; cell_n = bi / K + 1; }
                }
#pragma unroll
                for (int j = 0; j < LPL; ++j)
#pragma unroll
                    for (int n = 0; n < NV; ++n) cc[j][n] = c1[j][n];   // :137
                dt_int = dt_int + dt_red;                               // :138
            }
        } else if (a.method == MSED_RUNGE_KUTTA_4) {                    // :142-163 (stage formulas of column_kernel)
            double base[LPL][NV], c1[LPL][NV], a1[LPL][NV];
#define MSED_SPIN_ALL(stmt) _Pragma("unroll") for (int j = 0; j < LPL; ++j) _Pragma("unroll") for (int n = 0; n < NV; ++n) { stmt; }
            MSED_SPIN_ALL(base[j][n] = cc[j][n])
            rhs_of(cc, rhs);
            MSED_SPIN_ALL(c1[j][n] = fma(0.5 * dt, rhs[j][n], base[j][n]); a1[j][n] = 0.5 * rhs[j][n])
            rhs_of(c1, rhs);
            MSED_SPIN_ALL(c1[j][n] = fma(0.5 * dt, rhs[j][n], base[j][n]); a1[j][n] = a1[j][n] + rhs[j][n])
            rhs_of(c1, rhs);
            MSED_SPIN_ALL(c1[j][n] = fma(dt, rhs[j][n], base[j][n]); a1[j][n] = a1[j][n] + rhs[j][n])
            rhs_of(c1, rhs);
            MSED_SPIN_ALL(cc[j][n] = fma(dt * third, fma(0.5, rhs[j][n], a1[j][n]), base[j][n]))
            rhs_evals += 4;
        } else {                                                        // RK4 3/8, :164-185
            double base[LPL][NV], c1[LPL][NV], a1[LPL][NV], a2[LPL][NV];
            MSED_SPIN_ALL(base[j][n] = cc[j][n])
            rhs_of(cc, rhs);
            MSED_SPIN_ALL(c1[j][n] = fma(third * dt, rhs[j][n], base[j][n]); a1[j][n] = rhs[j][n])
            rhs_of(c1, rhs);
            MSED_SPIN_ALL(const double r0 = a1[j][n]; c1[j][n] = fma(dt, fma(-third, r0, rhs[j][n]), base[j][n]);
                          a1[j][n] = r0 - rhs[j][n]; a2[j][n] = fma(3.0, rhs[j][n], r0))
            rhs_of(c1, rhs);
            MSED_SPIN_ALL(c1[j][n] = fma(dt, a1[j][n] + rhs[j][n], base[j][n]); a2[j][n] = fma(3.0, rhs[j][n], a2[j][n]))
            rhs_of(c1, rhs);
            MSED_SPIN_ALL(cc[j][n] = fma(dt * 1.0 / 8.0, a2[j][n] + rhs[j][n], base[j][n]))
#undef MSED_SPIN_ALL
            rhs_evals += 4;
        }
    }

#pragma unroll
    for (int j = 0; j < LPL; ++j)
        if (active[j]) {
            double *out = p.buf[0] + (size_t)k[j] * ld + col;
#pragma unroll
            for (int n = 0; n < NV; ++n) out[(size_t)n * plane] = cc[j][n];
        }
    if (top) {
        if (a.nsteps > 0) {
#pragma unroll
            for (int n = NPART; n < NV; ++n) p.fluxes[(size_t)n * ld + col] = Ftop[n];   // driver :692
        }
        a.last_min_dt[col] = last_min_dt;
        a.grid_cell[4 * col + 0] = cell_k < 0 ? -99 : 1;
        a.grid_cell[4 * col + 1] = cell_k < 0 ? -99 : 1;
        a.grid_cell[4 * col + 2] = cell_k;
        a.grid_cell[4 * col + 3] = cell_n;
        a.counters[2 * col + 0] = rhs_evals;
        a.counters[2 * col + 1] = subcycles;
    }
}
