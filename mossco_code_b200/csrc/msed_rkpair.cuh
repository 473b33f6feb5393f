// msed_rkpair.cuh -- included inside namespace msed after msed_pair.cuh.
//
// Runge-Kutta stages chained in pairs (solver_library.F90:142-185).  The staged path needs four launches
// per step, each of which reads the stage state plus the base state and accumulators and writes them
// back (17 state passes per RK4 step).  The same one-layer-lag chaining as pair_kernel lets stage s+1
// run one layer behind stage s inside the thread, with the intermediate stage state in the shared-memory
// window: stages 1+2 and stages 3+4 become one launch each and an RK4 step costs 7 state passes
// (pair 12: read c, write c1 and acc; pair 34: read c1, c, acc, write c).  The stage formulas are those
// of the staged column_kernel ops, evaluated with the same inline arithmetic: both paths give
// bit-identical results (tests/test_gpu_fusion.py).
//
// Same scope as pair_kernel: no distributed POM flux cascade, profile != 3, closed-form porosity.
// RK has no accept test, so nothing here is speculative; check_NaN / clip act on the final stage only.

enum RkPair : int { RK4_12 = 0, RK4_34, RK38_12, RK38_34 };

// Shared memory of one CTA: a prefetch ring whose stage holds everything stage A reads of one layer
// (the state it evaluates; for stages 3+4 also the base state and the accumulators), and a two-layer
// window for the state handed from stage A to stage B.
constexpr int RK_WIN = 2;
__host__ __device__ constexpr int rk_ring_rows(int pair) { return (pair == RK4_12 || pair == RK38_12) ? NV : (pair == RK4_34 ? 3 * NV : 4 * NV); }
__host__ __device__ constexpr int rk_ring_stages(int pair) { return pair == RK38_34 ? 3 : 4; }   // 2 CTAs/SM must fit in 227 KB
__host__ __device__ constexpr size_t rk_pair_smem(int pair)
{
    return (size_t)(rk_ring_rows(pair) * rk_ring_stages(pair) + RK_WIN * NV) * ROW_BYTES;
}

#ifndef MSED_RKPAIR_MIN_BLOCKS
#define MSED_RKPAIR_MIN_BLOCKS 2
#endif

#ifdef MSED_RKPAIR_UNROLL
#define MSED_RKPAIR_UNROLL_PRAGMA MSED_UNROLL_PRAGMA(MSED_RKPAIR_UNROLL)
#else
#define MSED_RKPAIR_UNROLL_PRAGMA
#endif

template <int MODEL, int PAIR>
__global__ void __launch_bounds__(COL_BLOCK, MSED_RKPAIR_MIN_BLOCKS)
rk_pair_kernel(const __grid_constant__ KParams p)
{
    extern __shared__ __align__(16) double ring[];
    constexpr bool FIRST = (PAIR == RK4_12 || PAIR == RK38_12);   // stages 1+2, else stages 3+4
    constexpr bool IS38 = (PAIR == RK38_12 || PAIR == RK38_34);
    const Ctl *ctl = p.ctl;
    if (ctl->stop || ctl->steps_done >= ctl->steps_target) return;
    const int cur = ctl->cur;
    const bool do_clip = ctl->do_clip != 0;
    const double dt = ctl->dt;
    const double third = 1.0 / 3.0;
    // The stage-3 state normally never leaves shared memory.  The FABM diagnostics (msed_get_field)
    // describe the last get_rhs call, i.e. that state, so the last step of a call stores it where the
    // staged path leaves it (the spare buffer).
    const bool keep_c1 = !FIRST && (ctl->steps_done + 1 >= ctl->steps_target);

    const int col = p.col0 + blockIdx.x * COL_BLOCK + threadIdx.x;
    if (col >= p.col_end) return;
    if (p.mask[col] != 0) return;

    const int K = p.K;
    const size_t ld = p.ld;
    const size_t plane = (size_t)K * ld;
    // stage A reads conc (pair 12) or the stage state c1 left by pair 12 (pair 34)
    const double *in = (FIRST ? p.buf[cur] : p.buf[1 - cur]) + col;
    const double *base = p.buf[cur] + col;
    double *out = (FIRST ? p.buf[1 - cur] : p.buf[cur]) + col;   // c1, or conc itself for the final stage
    double *aux1 = p.aux1 + col, *aux2 = p.aux2 + col;

    constexpr int STAGES = rk_ring_stages(PAIR);
    constexpr uint32_t STAGE_B = (uint32_t)rk_ring_rows(PAIR) * ROW_BYTES;
    constexpr uint32_t WIN_B = NV * ROW_BYTES;
    const uint32_t sbase = smem_u32(ring) + threadIdx.x * 8u;
    const uint32_t wbase = sbase + STAGES * STAGE_B;
    const double *g_in = in, *g_fb = base, *g_f1 = aux1, *g_f2 = aux2;   // next layer to fetch
    int k_fetch = 0;
    auto fetch_next = [&]() {
        if (k_fetch < K) {
            const uint32_t sa = sbase + (uint32_t)(k_fetch % STAGES) * STAGE_B;
            const double *g = g_in, *gb = g_fb, *g1 = g_f1, *g2 = g_f2;
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                cp_async8(sa + n * ROW_BYTES, g);
                if (!FIRST) {
                    cp_async8(sa + (NV + n) * ROW_BYTES, gb);
                    cp_async8(sa + (2 * NV + n) * ROW_BYTES, g1);
                    if (IS38) cp_async8(sa + (3 * NV + n) * ROW_BYTES, g2);
                }
                g += plane; gb += plane; g1 += plane; g2 += plane;
            }
            g_in += ld; g_fb += ld; g_f1 += ld; g_f2 += ld;
        }
        ++k_fetch;
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) fetch_next();

    const double por_surf = (p.por_mode == 2) ? ld_ro(p.por + col) : 1.0;
    auto por_at = [&](int kk) -> double { return __dmul_rn(por_surf, p.portab[kk]); };
    const double temp = ld_ro(p.bdys + col);
    double cpart, cdiss, fT;
    column_constants<MODEL, false>(p, temp, cpart, cdiss, fT);

    auto top_boundary = [&](auto c0, double por0, double (&F)[NV], bool write_fluxes) {
        double Dp, Dd;
        top_coeffs(cpart, cdiss, por0, p.bf[0], Dp, Dd);
        const double rdz0 = 1.0 / p.dz[0];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const bool part = n < NPART;
            const int bc = part ? p.bcup_part : p.bcup_diss;
            double f = 0.0;
            if (bc == 1 || bc == 4) {
                f = ld_ro(p.fluxes + (size_t)n * ld + col);
            } else if (bc == 2) {
                const double Cup = ld_ro(p.bdys + (size_t)(n + 1) * ld + col);
                const double c1 = c0(n);
                const double C1 = part ? MSED_MUL(c1, por0) : c1;
                f = top_flux_dirichlet(part ? Dp : Dd, C1, Cup, rdz0);
            } else if (bc != 3 && n > 0) {
                f = F[n - 1];
            }
            F[n] = f;
            if (write_fluxes && !part) p.fluxes[(size_t)n * ld + col] = f;  // driver :692
        }
    };

    struct LayerCoef { double porc, porn, mDp, mDd, rpd; };
    auto make_coef = [&](auto has_next_tag, int kk) -> LayerCoef {
        LayerCoef lc;
        lc.porc = por_at(kk);
        lc.porn = lc.mDp = lc.mDd = 0.0;
        if (decltype(has_next_tag)::value) {
            lc.porn = por_at(kk + 1);
            interface_coeffs(cpart, cdiss, lc.porc, lc.porn, p.bf[kk + 1], p.rdzc[kk], lc.mDp, lc.mDd);
        }
        lc.rpd = fast_rcp(MSED_MUL(lc.porc, p.dz[kk]));
        return lc;
    };
    // right-hand side of one layer (same inline arithmetic as column_kernel)
    auto layer_rates = [&](auto has_next_tag, const LayerCoef &lc, const double (&cc)[NV], auto cn,
                           double (&F)[NV], double (&rhs)[NV]) {
        double Fn[NV];
        if (decltype(has_next_tag)::value) {
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const double c = cn(n);
                if (n < NPART) Fn[n] = flux_particulate(lc.mDp, c, lc.porn, cc[n], lc.porc);
                else Fn[n] = flux_dissolved(lc.mDd, c, cc[n]);
            }
        } else {
#pragma unroll
            for (int n = 0; n < NV; ++n) Fn[n] = 0.0;
        }
        double r[NV];
        if (MODEL == MSED_MODEL_OMEXDIA_P) {
            omexdia_rates(p.om, cc, fT, r, nullptr);
        } else {
#pragma unroll
            for (int n = 0; n < NV; ++n) r[n] = 0.0;
        }
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            rhs[n] = layer_rhs(F[n], Fn[n], lc.rpd, r[n]);
            F[n] = Fn[n];
        }
    };

    cp_async_wait<STAGES - 2>();
    double FA[NV], FB[NV];
    top_boundary([&](int n) { return lds64(sbase + n * ROW_BYTES); }, por_at(0), FA, false);

    // what stage B needs from stage A one iteration later: the base state of the layer and one vector
    //   RK4_12: x = 0.5*k1   RK4_34: x = acc + k3   RK38_12: x = k1   RK38_34: x = Q + 3*k3
    double carry_base[NV], carry_x[NV];
    LayerCoef coef_prev;
    bool nanf = false;
    double *g_out = out, *g_w1 = aux1, *g_w2 = aux2;   // write position of stage B (layer j)
    double *g_c1 = p.buf[1 - cur] + col;               // stage-3 state of layer k (keep_c1 only)

    // ---- stage A on layer k -----------------------------------------------------------------------
    auto stage_a = [&](auto has_next_tag, int k, double (&baseA)[NV], double (&xA)[NV]) -> LayerCoef {
        fetch_next();
        cp_async_wait<STAGES - 2>();
        const uint32_t sc = sbase + (uint32_t)(k % STAGES) * STAGE_B;
        const uint32_t sn = sbase + (uint32_t)((k + 1) % STAGES) * STAGE_B;
        const uint32_t wk = wbase + (uint32_t)(k & (RK_WIN - 1)) * WIN_B;
        double cc[NV], a1[NV], a2[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            cc[n] = lds64(sc + n * ROW_BYTES);
            if (FIRST) {
                baseA[n] = cc[n];                        // stage 1 evaluates the base state itself
            } else {
                baseA[n] = lds64(sc + (NV + n) * ROW_BYTES);
                a1[n] = lds64(sc + (2 * NV + n) * ROW_BYTES);
                if (IS38) a2[n] = lds64(sc + (3 * NV + n) * ROW_BYTES);
            }
        }
        const LayerCoef lc = make_coef(has_next_tag, k);
        double rhs[NV];
        layer_rates(has_next_tag, lc, cc, [&](int n) { return lds64(sn + n * ROW_BYTES); }, FA, rhs);
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            double yb;  // the state the next stage evaluates
            if (PAIR == RK4_12) {              // :147  c1 = c + 0.5*dt*k1 ; acc = 0.5*k1
                yb = fma(0.5 * dt, rhs[n], baseA[n]);
                xA[n] = 0.5 * rhs[n];
            } else if (PAIR == RK4_34) {       // :156  c1 = c + dt*k3 ; acc += k3
                yb = fma(dt, rhs[n], baseA[n]);
                xA[n] = a1[n] + rhs[n];
            } else if (PAIR == RK38_12) {      // :169  c1 = c + third*dt*k1
                yb = fma(third * dt, rhs[n], baseA[n]);
                xA[n] = rhs[n];
            } else {                           // :178  c1 = c + dt*(P + k3) ; Q += 3*k3
                yb = fma(dt, a1[n] + rhs[n], baseA[n]);
                xA[n] = fma(3.0, rhs[n], a2[n]);
            }
            sts64(wk + n * ROW_BYTES, yb);
            if (!FIRST && keep_c1) g_c1[(size_t)n * plane] = yb;
        }
        g_c1 += ld;
        return lc;
    };

    // ---- stage B on layer j (one layer behind) ------------------------------------------------------
    auto stage_b = [&](auto has_next_tag, auto clip_tag, int j, const LayerCoef &lc, const double (&baseB)[NV],
                       const double (&xB)[NV]) {
        const uint32_t wj = wbase + (uint32_t)(j & (RK_WIN - 1)) * WIN_B;
        const uint32_t wn = wbase + (uint32_t)((j + 1) & (RK_WIN - 1)) * WIN_B;
        if (j == 0) top_boundary([&](int n) { return lds64(wj + n * ROW_BYTES); }, por_at(0), FB, true);
        double cc[NV], rhs[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) cc[n] = lds64(wj + n * ROW_BYTES);
        layer_rates(has_next_tag, lc, cc, [&](int n) { return lds64(wn + n * ROW_BYTES); }, FB, rhs);
        double *go = g_out, *gw1 = g_w1, *gw2 = g_w2;
        double raw[NV];  // new state before the clip
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            double newc;
            if (PAIR == RK4_12) {              // :152  c1 = c + 0.5*dt*k2 ; acc = 0.5*k1 + k2
                newc = fma(0.5 * dt, rhs[n], baseB[n]);
                *gw1 = xB[n] + rhs[n];
            } else if (PAIR == RK4_34) {       // :160  c = c + dt*third*(acc + 0.5*k4)
                newc = fma(dt * third, fma(0.5, rhs[n], xB[n]), baseB[n]);
            } else if (PAIR == RK38_12) {      // :174  c1 = c + dt*(k2 - third*k1) ; P = k1-k2 ; Q = k1+3*k2
                newc = fma(dt, fma(-third, xB[n], rhs[n]), baseB[n]);
                *gw1 = xB[n] - rhs[n];
                *gw2 = fma(3.0, rhs[n], xB[n]);
            } else {                           // :182  c = c + dt*1/8*(Q' + k4)
                newc = fma(dt * 1.0 / 8.0, xB[n] + rhs[n], baseB[n]);
            }
            raw[n] = newc;
            if (!FIRST && decltype(clip_tag)::value) {  // final stage: check_NaN + clip (component :1718-1732)
                if (n & 1) nanf |= either_nan(raw[n - 1], raw[n]);
                const double mn = p.om.minimum[n];
                newc = clip_min(newc, mn);
            }
            *go = newc;
            go += plane; gw1 += plane; gw2 += plane;
        }
        g_out += ld; g_w1 += ld; g_w2 += ld;
    };

    auto sweep = [&](auto clip_tag) {
        using Y = std::true_type;
        using N = std::false_type;
        if (K == 1) {
            coef_prev = stage_a(N{}, 0, carry_base, carry_x);
            stage_b(N{}, clip_tag, 0, coef_prev, carry_base, carry_x);
            return;
        }
        coef_prev = stage_a(Y{}, 0, carry_base, carry_x);
        MSED_RKPAIR_UNROLL_PRAGMA
        for (int k = 1; k < K - 1; ++k) {  // steady state: stage A on layer k, stage B on layer k-1
            double nb[NV], nx[NV];
            const LayerCoef lc = stage_a(Y{}, k, nb, nx);
            stage_b(Y{}, clip_tag, k - 1, coef_prev, carry_base, carry_x);
            coef_prev = lc;
#pragma unroll
            for (int n = 0; n < NV; ++n) { carry_base[n] = nb[n]; carry_x[n] = nx[n]; }
        }
        double nb[NV], nx[NV];
        const LayerCoef last = stage_a(N{}, K - 1, nb, nx);
        stage_b(Y{}, clip_tag, K - 2, coef_prev, carry_base, carry_x);
        stage_b(N{}, clip_tag, K - 1, last, nb, nx);
    };
    if (!FIRST && do_clip) sweep(std::true_type{});
    else sweep(std::false_type{});
    cp_async_wait<0>();

    if (!FIRST && nanf) atomicOr(&p.ctl->flags[1], 1);
}
