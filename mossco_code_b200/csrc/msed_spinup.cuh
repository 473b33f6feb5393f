// msed_spinup.cuh -- included inside namespace msed after msed_column.cuh.
//
// The 1-D pre-simulation of the component (fabm_sediment_component.F90:557-632) for a whole BATCH of
// independent columns in one launch: one warp per member, lane k = layer k (knum <= 32), the column in
// registers for all nsteps ode_solver calls (dt_spinup = 3600 s, :574).  A member is what the reference calls
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


template <int MODEL>
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
    const bool active = lane < K;
    const int k = active ? lane : K - 1;   // spare lanes shadow the deepest layer; they never store
    const bool has_next = lane + 1 < K;
    const bool top = lane == 0;
    const size_t ld = p.ld;
    const size_t plane = (size_t)K * ld;

    double cc[NV];
    {
        const double *in = p.buf[0] + (size_t)k * ld + col;
#pragma unroll
        for (int n = 0; n < NV; ++n) cc[n] = in[(size_t)n * plane];
    }
    // step-invariant coefficients (as chain_kernel)
    const double por_surf = (p.por_mode == 2) ? ld_ro(p.por + col) : 1.0;
    const double temp = ld_ro(p.bdys + col);
    double cpart, cdiss, fT;
    column_constants<MODEL, false>(p, temp, cpart, cdiss, fT);
    const double porc = __dmul_rn(por_surf, p.portab[k]);
    double porn = 0.0, mDp = 0.0, mDd = 0.0;
    if (has_next) {
        porn = __dmul_rn(por_surf, p.portab[k + 1]);
        interface_coeffs(cpart, cdiss, porc, porn, p.bf[k + 1], p.rdzc[k], mDp, mDd);
    }
    const double rpd = fast_rcp(MSED_MUL(porc, p.dz[k]));
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
    // get_rhs for the state x of this lane's layer
    auto rhs_of = [&](const double (&x)[NV], double (&rhs)[NV]) {
        double cn[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) cn[n] = __shfl_down_sync(FULL, x[n], 1);
        double r[NV];
        if (MODEL == MSED_MODEL_OMEXDIA_P) {
            omexdia_rates(om, x, fT, r, nullptr);
        } else {
#pragma unroll
            for (int n = 0; n < NV; ++n) r[n] = 0.0;
        }
        double Fn[NV], F[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const double f = (n < NPART) ? flux_particulate(mDp, cn[n], porn, x[n], porc) : flux_dissolved(mDd, cn[n], x[n]);
            Fn[n] = has_next ? f : 0.0;
        }
#pragma unroll
        for (int n = 0; n < NV; ++n) F[n] = __shfl_up_sync(FULL, Fn[n], 1);
#pragma unroll
        for (int n = 0; n < NPART; ++n) F[n] = top ? tin[n] : F[n];
        if (bc_diss == 2) {
#pragma unroll
            for (int n = NPART; n < NV; ++n) {
                const double f = top_flux_dirichlet(Dd0, x[n], tin[n], rdz0);
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
        for (int n = 0; n < NV; ++n) {
            Ftop[n] = F[n];
            rhs[n] = layer_rhs(F[n], Fn[n], rpd, r[n]);
        }
    };

    double last_min_dt = a.last_min_dt0;
    int cell_k = -99, cell_n = -99;
    long long rhs_evals = 0, subcycles = 0;
    const double dt = a.dt, third = 1.0 / 3.0;

    for (long long s = 0; s < a.nsteps; ++s) {
        double rhs[NV];
        if (a.method == MSED_EULER) {                                   // solver_library.F90:99-102
            rhs_of(cc, rhs);
#pragma unroll
            for (int n = 0; n < NV; ++n) cc[n] = euler_update(dt, rhs[n], cc[n]);
            rhs_evals += 1;
        } else if (a.method == MSED_ADAPTIVE_EULER) {                   // :104-140
            double dt_int = 0.0, dt_red = dt;
            while (dt_int < dt) {
                rhs_of(cc, rhs);
                rhs_evals += 1;
                double c1[NV];
                bool viol = false;
#pragma unroll
                for (int n = 0; n < NV; ++n) {
                    c1[n] = euler_update(dt_red, rhs[n], cc[n]);
                    viol |= violates(p.fac, cc[n], c1[n]);              // :121
                }
                if (__any_sync(FULL, viol && active) && dt_red > a.dt_min) {   // :126-128
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
                    for (int n = 0; n < NV; ++n) {
                        const double v = __ddiv_rn(__dsub_rn(c1[n], cc[n]), cc[n]);
                        const int idx = k + K * n;
                        if (active && v == v && (bi == 0x7fffffff || v < bv)) { bv = v; bi = idx; }
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        const double ov = __shfl_xor_sync(FULL, bv, o);
                        const int oi = __shfl_xor_sync(FULL, bi, o);
                        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
                    }
                    if (bi != 0x7fffffff) { cell_k = bi % K + 1; cell_n = bi / K + 1; }
                }
#pragma unroll
                for (int n = 0; n < NV; ++n) cc[n] = c1[n];             // :137
                dt_int = dt_int + dt_red;                               // :138
            }
        } else if (a.method == MSED_RUNGE_KUTTA_4) {                    // :142-163 (stage formulas of column_kernel)
            double base[NV], c1[NV], a1[NV];
#pragma unroll
            for (int n = 0; n < NV; ++n) base[n] = cc[n];
            rhs_of(cc, rhs);
#pragma unroll
            for (int n = 0; n < NV; ++n) { c1[n] = fma(0.5 * dt, rhs[n], base[n]); a1[n] = 0.5 * rhs[n]; }
            rhs_of(c1, rhs);
#pragma unroll
            for (int n = 0; n < NV; ++n) { c1[n] = fma(0.5 * dt, rhs[n], base[n]); a1[n] = a1[n] + rhs[n]; }
            rhs_of(c1, rhs);
#pragma unroll
            for (int n = 0; n < NV; ++n) { c1[n] = fma(dt, rhs[n], base[n]); a1[n] = a1[n] + rhs[n]; }
            rhs_of(c1, rhs);
#pragma unroll
            for (int n = 0; n < NV; ++n) cc[n] = fma(dt * third, fma(0.5, rhs[n], a1[n]), base[n]);
            rhs_evals += 4;
        } else {                                                        // RK4 3/8, :164-185
            double base[NV], c1[NV], a1[NV], a2[NV];
#pragma unroll
            for (int n = 0; n < NV; ++n) base[n] = cc[n];
            rhs_of(cc, rhs);
#pragma unroll
            for (int n = 0; n < NV; ++n) { c1[n] = fma(third * dt, rhs[n], base[n]); a1[n] = rhs[n]; }
            rhs_of(c1, rhs);
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const double r0 = a1[n];
                c1[n] = fma(dt, fma(-third, r0, rhs[n]), base[n]);
                a1[n] = r0 - rhs[n];
                a2[n] = fma(3.0, rhs[n], r0);
            }
            rhs_of(c1, rhs);
#pragma unroll
            for (int n = 0; n < NV; ++n) { c1[n] = fma(dt, a1[n] + rhs[n], base[n]); a2[n] = fma(3.0, rhs[n], a2[n]); }
            rhs_of(c1, rhs);
#pragma unroll
            for (int n = 0; n < NV; ++n) cc[n] = fma(dt * 1.0 / 8.0, a2[n] + rhs[n], base[n]);
            rhs_evals += 4;
        }
    }

    if (active) {
        double *out = p.buf[0] + (size_t)k * ld + col;
#pragma unroll
        for (int n = 0; n < NV; ++n) out[(size_t)n * plane] = cc[n];
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
