// msed_rkquad.cuh -- included inside namespace msed after msed_pair.cuh.
//
// All four stages of a Runge-Kutta call (solver_library.F90:142-185) in ONE pass over HBM, thread per column.
//
// rk_pair_kernel chains the stages two by two and an RK4 call costs seven state passes; on a tile too large for
// a warp per column (rk_chain_kernel) those two launches run at the HBM roofline of the bytes they move and the
// fp64 pipe idles.  Here the one-layer-lag chaining of pair_kernel is taken to four stages: in iteration k of
// the walk down the column
//     stage 1 evaluates layer k   of the state c               (read from the input ring),
//     stage 2 evaluates layer k-1 of  c + a21 dt k1            (link slot 2 + stage 1's fresh layer k),
//     stage 3 evaluates layer k-2 of the third stage state     (link slot 3 + stage 2's fresh layer k-1),
//     stage 4 evaluates layer k-3 of the fourth stage state    (link slot 4 + stage 3's fresh layer k-2)
//             and stores the new state of layer k-3 -- in place: the ring has long fetched that layer.
// A stage hands the layer it has just produced to the next stage in registers (as that stage's "layer below")
// and through a one-layer link slot in shared memory (as the next iteration's "current layer"); the weighted
// sums of the k_i go from stage to stage one iteration at a time (the first in registers, the others through shared
// memory, read where they are consumed); the base state of stages
// 2-4 is still in the input ring (6 slots: layers k-3 .. k+1 in use, k+2 in flight).  The state is read
// once and written once per CALL (2 passes instead of 7), nothing else touches HBM.
//
// The stage formulas and the inline arithmetic are those of rk_pair_kernel / the staged column_kernel ops, so
// the three paths give bit-identical results (tests/test_gpu_fusion.py).  Same scope as rk_pair_kernel (no
// distributed POM flux cascade, profile != 3, closed-form porosity), and knum >= 5 (the peeled head and tail
// of the walk); anything else stays with the stage pairs.
//
// Shared memory: (6 ring + 3 link + 2 k-sum) layers x 8 variables x 128 columns x 8 B = 88 KB per CTA, plus 9 KB of
// layer coefficients (three per layer, made by stage 1, read by the others; RK4-3/8 keeps its third k sum there
// instead): two CTAs per SM fit the 200 KB carve-out that leaves the L1 its 56 KB; 255 registers per thread: eight
// warps per SM, each with four independent RHS evaluations in flight.

constexpr int RKQ_RING = 6;                                   // input ring slots: layers k-3 .. k+1 in use, k+2 in flight
constexpr uint32_t RKQ_STAGE_B = NV * ROW_BYTES;              // one layer of one CTA
constexpr int RKQ_COEF_ROWS = 3;                              // LayerCoef: mDp, mDd, rpd
constexpr int RKQ_COEF_SLOTS = 3;                             // layers k-2 .. k (k-3 is read before k takes its slot)
// ring, three link slots, two weighted sums of the k_i, layer coefficients
constexpr size_t RKQ_SMEM_BYTES = (size_t)(RKQ_RING + 3 + 2) * RKQ_STAGE_B + RKQ_COEF_SLOTS * RKQ_COEF_ROWS * ROW_BYTES;
constexpr int RKQ_MIN_LAYERS = 5;

#ifndef MSED_RKQUAD_MIN_BLOCKS
#define MSED_RKQUAD_MIN_BLOCKS 2
#endif

template <int MODEL, bool IS38>
__global__ void __launch_bounds__(COL_BLOCK, MSED_RKQUAD_MIN_BLOCKS)
rk_quad_kernel(const __grid_constant__ KParams p)
{
    extern __shared__ __align__(16) double ring[];
    const Ctl *ctl = p.ctl;
    if (ctl->stop || ctl->steps_done >= ctl->steps_target) return;
    const int cur = ctl->cur;
    const bool do_clip = ctl->do_clip != 0;
    const double dt = ctl->dt;
    const double third = 1.0 / 3.0;
    // the FABM diagnostics (msed_get_field) describe the last get_rhs call, i.e. the stage-4 input state: the last
    // call of a sequence stores it where the staged path leaves it (the spare buffer)
    const bool keep_c1 = ctl->steps_done + 1 >= ctl->steps_target;

    // On a tile with land the launch runs over the list of wet columns (KParams::colmap, as pair_kernel does): a CTA
    // of mostly land columns would hold its 97 KB of shared memory for a few wet warps.
    const int t = p.col0 + blockIdx.x * COL_BLOCK + threadIdx.x;
    if (t >= p.col_end) return;
    const int col = p.colmap ? p.colmap[t] : t;
    if (!p.colmap && p.mask[col] != 0) return;   // (the list holds wet columns only: msed_set_mask builds both)

    const int K = p.K;
    const size_t ld = p.ld;
    const size_t plane = (size_t)K * ld;
    double *conc = p.buf[cur] + col;          // read through the ring, rewritten three layers behind
    double *g_c1 = p.buf[1 - cur] + col;      // keep_c1 only
    double *g_out = conc;

    const uint32_t sbase = smem_u32(ring) + threadIdx.x * 8u;
    auto q_lds = [&](uint32_t addr) __attribute__((always_inline)) -> double { return lds64(addr); };
    auto q_sts = [&](uint32_t addr, double v) __attribute__((always_inline)) { sts64(addr, v); };
    auto stg64 = [](double *g, double v) __attribute__((always_inline)) {
        asm volatile("st.global.f64 [%0], %1;" ::"l"(g), "d"(v) : "memory");
    };
    const uint32_t lk2 = sbase + RKQ_RING * RKQ_STAGE_B, lk3 = lk2 + RKQ_STAGE_B, lk4 = lk3 + RKQ_STAGE_B;
    const uint32_t xsa = lk4 + RKQ_STAGE_B;   // x34 of the layer stage 4 evaluates next
    const uint32_t xsb = xsa + RKQ_STAGE_B;   // x23 (RK4) / x23b (RK4-3/8) of the layer stage 3 evaluates next
    const uint32_t cfb = xsb + RKQ_STAGE_B;   // layer coefficients: slot (layer mod 3) x {mDp, mDd, rpd}
    int kmod = 0;                              // k mod RKQ_RING (the ring is not a power of two: counted, not divided)
    auto slot = [&](int rel) __attribute__((always_inline)) -> uint32_t {   // ring slot of layer k + rel, rel = -3 .. 2
        int sidx = kmod + rel;
        if (sidx < 0) sidx += RKQ_RING;
        if (sidx >= RKQ_RING) sidx -= RKQ_RING;
        return sbase + (uint32_t)sidx * RKQ_STAGE_B;
    };
    const double *g_in = conc;
    int k_fetch = 0, fmod = 0;
    auto fetch_next = [&]() __attribute__((always_inline)) {
        if (k_fetch < K) {
            const uint32_t sa = sbase + (uint32_t)fmod * RKQ_STAGE_B;
            const double *g = g_in;
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                cp_async8(sa + n * ROW_BYTES, g);
                g += plane;
            }
            g_in += ld;
        }
        ++k_fetch;
        fmod = (fmod == RKQ_RING - 1) ? 0 : fmod + 1;
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < RKQ_RING - 4; ++s) fetch_next();      // layers 0 and 1

    const double por_surf = (p.por_mode == 2) ? ld_ro(p.por + col) : 1.0;
    auto por_at = [&](int kk) __attribute__((always_inline)) -> double { return __dmul_rn(por_surf, p.portab[kk]); };
    const double temp = ld_ro(p.bdys + col);
    // the upper-boundary value of every variable (an input flux or a concentration above the bed): every stage needs
    // them once, in the first four iterations -- read here in one go, beside the temperature, they cost one trip to
    // L2 / HBM per column instead of one per stage and variable at the head of the walk
    double ub[NV];
#pragma unroll
    for (int n = 0; n < NV; ++n) {
        const int bc = (n < NPART) ? p.bcup_part : p.bcup_diss;
        ub[n] = 0.0;
        if (bc == 1 || bc == 4) ub[n] = ld_ro(p.fluxes + (size_t)n * ld + col);          // :783,:792
        else if (bc == 2) ub[n] = ld_ro(p.bdys + (size_t)(n + 1) * ld + col);            // :786
    }
    double cpart, cdiss, fT;
    column_constants<MODEL, false>(p, temp, cpart, cdiss, fT);

    // upper boundary of one stage: F[n] = Flux(1) (diff3d :782-803) from the stage's own layer-1 state
    auto top_boundary = [&](const double (&c0)[NV], double (&F)[NV], bool write_fluxes) __attribute__((always_inline)) {
        double Dp, Dd;
        const double por0 = por_at(0);
        top_coeffs(cpart, cdiss, por0, p.bf[0], Dp, Dd);
        const double rdz0 = 1.0 / p.dz[0];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const bool part = n < NPART;
            const int bc = part ? p.bcup_part : p.bcup_diss;
            double f = 0.0;
            if (bc == 1 || bc == 4) {
                f = ub[n];
            } else if (bc == 2) {
                const double Cup = ub[n];
                const double C1 = part ? MSED_MUL(c0[n], por0) : c0[n];
                f = top_flux_dirichlet(part ? Dp : Dd, C1, Cup, rdz0);
            } else if (bc != 3 && n > 0) {
                f = F[n - 1];
            }
            F[n] = f;
            if (write_fluxes && !part) p.fluxes[(size_t)n * ld + col] = f;  // driver :692
        }
    };

    // state-independent coefficients of a layer and of its lower interface (every stage needs the same ones, one
    // iteration after the other): -D/dzc of the interface for particulates / solutes, 1/(porosity*dz)
    struct LayerCoef { double mDp, mDd, rpd; };
    auto make_coef = [&](auto has_next_tag, int kk, double porc, double porn) __attribute__((always_inline)) -> LayerCoef {
        LayerCoef lc;
        lc.mDp = lc.mDd = 0.0;
        if (decltype(has_next_tag)::value)
            interface_coeffs(cpart, cdiss, porc, porn, p.bf[kk + 1], p.rdzc[kk], lc.mDp, lc.mDd);
        lc.rpd = fast_rcp(MSED_MUL(porc, p.dz[kk]));
        return lc;
    };
    // right-hand side of one layer (the inline arithmetic of column_kernel)
    auto layer_rates = [&](auto has_next_tag, const LayerCoef &lc, double porc, double porn, const double (&cc)[NV],
                           const double (&cn)[NV], double (&F)[NV], double (&rhs)[NV]) __attribute__((always_inline)) {
        double Fn[NV];
        if (decltype(has_next_tag)::value) {
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                if (n < NPART) Fn[n] = flux_particulate(lc.mDp, cn[n], porn, cc[n], porc);
                else Fn[n] = flux_dissolved(lc.mDd, cn[n], cc[n]);
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

    // ---- what travels from iteration to iteration in registers ---------------------------------------------
    double F1[NV], F2[NV], F3[NV], F4[NV];     // flux through the upper interface of each stage's next layer
    // weighted sums of the k_i, handed from stage to stage (the layer a stage finishes in iteration k is the
    // next stage's layer in iteration k+1):
    //   RK4   x12 = 0.5 k1      x23 = 0.5 k1 + k2             x34 = x23 + k3                (:147-160)
    //   RK38  x12 = k1          x23 = P = k1 - k2, x23b = Q = k1 + 3 k2     x34 = Q + 3 k3  (:169-182)
    // Each is consumed by its stage before the stage in front of it writes the next layer's value into the same
    // registers (the updates sit at the end of the iteration, last stage first): no copies.
    double x12[NV], x23[NV];   // (x34 and x23 / x23b live in shared memory: see xsa, xsb)
    // RK4-3/8 carries one sum more than RK4 (P and Q beside k1): with the coefficient ring it spilled in its loop.  It
    // gives the ring's shared memory to P instead and makes the coefficients of a layer anew in every stage (14 fp64
    // instructions per stage and layer: cheaper than the reloads of the spills).
    constexpr bool COEF_RING = !IS38;
    const uint32_t xsp = cfb;                  // IS38: x23 = P of the layer stage 3 evaluates next
    // The state-independent coefficients of a layer are made once, by stage 1, and wait for the other stages in a
    // three-slot ring in shared memory (as registers they were 18 of the 255, and the kernel spilled in its loop).
    int kmod3 = 0;                             // k mod 3
    auto coef_addr = [&](int back) __attribute__((always_inline)) -> uint32_t {   // slot of layer k - back, back = 0..3
        int sidx = kmod3 - (back % 3);
        if (sidx < 0) sidx += 3;
        return cfb + (uint32_t)sidx * (RKQ_COEF_ROWS * ROW_BYTES);
    };
    auto coef_load = [&](int back) __attribute__((always_inline)) -> LayerCoef {
        const uint32_t a = coef_addr(back);
        LayerCoef lc;
        lc.mDp = lds64(a);
        lc.mDd = lds64(a + ROW_BYTES);
        lc.rpd = lds64(a + 2 * ROW_BYTES);
        return lc;
    };
#pragma unroll
    for (int n = 0; n < NV; ++n) F1[n] = F2[n] = F3[n] = F4[n] = x12[n] = x23[n] = 0.0;
    bool nanf = false;

    // One iteration of the walk.  Mk: what stage k does in it -- 0 nothing, 1 the first layer of the column (upper
    // boundary first), 2 an inner layer, 3 the last layer (closed bottom).  Stage s works on layer k - (s-1).
    auto iteration = [&](auto m1, auto m2, auto m3, auto m4, int k) __attribute__((always_inline)) {
        constexpr int M1 = decltype(m1)::value, M2 = decltype(m2)::value, M3 = decltype(m3)::value, M4 = decltype(m4)::value;
        using Y = std::true_type;
        using N = std::false_type;
        // the link slots hold what the previous iteration produced: read them before this iteration's results
        // take their place
        double c2[NV], c3[NV], c4[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            if (M2) c2[n] = q_lds(lk2 + n * ROW_BYTES);
            if (M3) c3[n] = q_lds(lk3 + n * ROW_BYTES);
            if (M4) c4[n] = q_lds(lk4 + n * ROW_BYTES);
        }
        double y2n[NV], y3n[NV], y4n[NV];      // the layer stages 1-3 produce in this iteration
        double rhs1[NV], rhs2[NV], rhs3[NV];
        LayerCoef cf4;
        if (M4 && COEF_RING) cf4 = coef_load(3);   // (its slot is the one stage 1 fills below)
        if (M1) {                              // ---- stage 1, layer k: k1 = f(c) ------------------------------
            fetch_next();                      // layer k+2 into the slot layer k-4 has left
            cp_async_wait<RKQ_RING - 5>();     // layer k+1 has landed
            const uint32_t sc = slot(0), sn = slot(1);
            double cc[NV], cn[NV];
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                cc[n] = q_lds(sc + n * ROW_BYTES);
                cn[n] = (M1 != 3) ? q_lds(sn + n * ROW_BYTES) : 0.0;
            }
            if (M1 == 1) top_boundary(cc, F1, false);
            const double pk = por_at(k), pk1 = (M1 != 3) ? por_at(k + 1) : 0.0;
            LayerCoef cf1;
            if (M1 != 3) { cf1 = make_coef(Y{}, k, pk, pk1); layer_rates(Y{}, cf1, pk, pk1, cc, cn, F1, rhs1); }
            else         { cf1 = make_coef(N{}, k, pk, pk1); layer_rates(N{}, cf1, pk, pk1, cc, cn, F1, rhs1); }
            if (COEF_RING) {
                const uint32_t a = coef_addr(0);
                q_sts(a, cf1.mDp);
                q_sts(a + ROW_BYTES, cf1.mDd);
                q_sts(a + 2 * ROW_BYTES, cf1.rpd);
            }
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                if (!IS38) y2n[n] = fma(0.5 * dt, rhs1[n], cc[n]);      // :147  c1 = c + 0.5*dt*k1
                else       y2n[n] = fma(third * dt, rhs1[n], cc[n]);    // :169  c1 = c + third*dt*k1
                q_sts(lk2 + n * ROW_BYTES, y2n[n]);
            }
        }
        if (M2) {                              // ---- stage 2, layer k-1: k2 = f(c1) ---------------------------
            const uint32_t sb = slot(-1);
            if (M2 == 1) top_boundary(c2, F2, false);
            const LayerCoef cf2 = COEF_RING ? coef_load(1)
                                            : (M2 != 3 ? make_coef(Y{}, k - 1, por_at(k - 1), por_at(k))
                                                       : make_coef(N{}, k - 1, por_at(k - 1), 0.0));
            double base[NV];   // (read in one go, ahead of their use: a load next to its use inside the ordered
#pragma unroll             //  ld.shared / st.shared sequence exposes its latency eight times per stage)
            for (int n = 0; n < NV; ++n) base[n] = q_lds(sb + n * ROW_BYTES);
            if (M2 != 3) layer_rates(Y{}, cf2, por_at(k - 1), por_at(k), c2, y2n, F2, rhs2);
            else         layer_rates(N{}, cf2, por_at(k - 1), 0.0, c2, c2, F2, rhs2);
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                if (!IS38) y3n[n] = fma(0.5 * dt, rhs2[n], base[n]);                     // :152  c1 = c + 0.5*dt*k2
                else       y3n[n] = fma(dt, fma(-third, x12[n], rhs2[n]), base[n]);      // :174  c1 = c + dt*(k2 - third*k1)
                q_sts(lk3 + n * ROW_BYTES, y3n[n]);
            }
        }
        if (M3) {                              // ---- stage 3, layer k-2: k3 = f(c1) ---------------------------
            const uint32_t sb = slot(-2);
            if (M3 == 1) top_boundary(c3, F3, false);
            const LayerCoef cf3 = COEF_RING ? coef_load(2)
                                            : (M3 != 3 ? make_coef(Y{}, k - 2, por_at(k - 2), por_at(k - 1))
                                                       : make_coef(N{}, k - 2, por_at(k - 2), 0.0));
            if (IS38) {
#pragma unroll
                for (int n = 0; n < NV; ++n) x23[n] = q_lds(xsp + n * ROW_BYTES);
            }
            double base[NV];
#pragma unroll
            for (int n = 0; n < NV; ++n) base[n] = q_lds(sb + n * ROW_BYTES);
            if (M3 != 3) layer_rates(Y{}, cf3, por_at(k - 2), por_at(k - 1), c3, y3n, F3, rhs3);
            else         layer_rates(N{}, cf3, por_at(k - 2), 0.0, c3, c3, F3, rhs3);
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                if (!IS38) y4n[n] = fma(dt, rhs3[n], base[n]);                           // :156  c1 = c + dt*k3
                else       y4n[n] = fma(dt, MSED_ADD(x23[n], rhs3[n]), base[n]);         // :178  c1 = c + dt*(P + k3)
                q_sts(lk4 + n * ROW_BYTES, y4n[n]);
            }
            if (keep_c1) {
                double *gc = g_c1;   // (pointer steps instead of n*plane offsets: those were eight more loop invariants in
#pragma unroll               //  a kernel that already spills its uniform registers)
                for (int n = 0; n < NV; ++n) {
                    stg64(gc, y4n[n]);
                    gc += plane;
                    asm volatile("" : "+l"(gc));
                }
            }
            g_c1 += ld;
        }
        if (M4) {                              // ---- stage 4, layer k-3: k4 = f(c1), the new state ------------
            const uint32_t sb = slot(-3);
            double rhs[NV], raw[NV], base[NV], x34[NV];
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                base[n] = q_lds(sb + n * ROW_BYTES);
                x34[n] = q_lds(xsa + n * ROW_BYTES);
            }
            if (!COEF_RING) cf4 = (M4 != 3) ? make_coef(Y{}, k - 3, por_at(k - 3), por_at(k - 2))
                                            : make_coef(N{}, k - 3, por_at(k - 3), 0.0);
            if (M4 == 1) top_boundary(c4, F4, true);
            if (M4 != 3) layer_rates(Y{}, cf4, por_at(k - 3), por_at(k - 2), c4, y4n, F4, rhs);
            else         layer_rates(N{}, cf4, por_at(k - 3), 0.0, c4, c4, F4, rhs);
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                if (!IS38) raw[n] = fma(dt * third, fma(0.5, rhs[n], x34[n]), base[n]);          // :160
                else       raw[n] = fma(dt * 1.0 / 8.0, MSED_ADD(x34[n], rhs[n]), base[n]);      // :182
            }
            if (do_clip) {                     // check_NaN + clip (component :1718-1732)
#pragma unroll
                for (int n = 0; n < NV; ++n) {
                    if (n & 1) nanf |= either_nan(raw[n - 1], raw[n]);
                    raw[n] = clip_min(raw[n], p.om.minimum[n]);
                }
            }
            double *go = g_out;
#pragma unroll
            for (int n = 0; n < NV; ++n) {   // (the empty asm keeps the steps: folded into n*plane offsets they are
                stg64(go, raw[n]);           // fourteen more uniform registers, which the kernel does not have)
                go += plane;
                asm volatile("" : "+l"(go));
            }
            g_out += ld;
        }
        // hand-over to the next iteration, last stage first (see x12 .. x34 above)
        double xb[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n)
            if (M3) xb[n] = q_lds(xsb + n * ROW_BYTES);
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            if (!IS38) {
                if (M3) q_sts(xsa + n * ROW_BYTES, MSED_ADD(xb[n], rhs3[n]));                        // :156  acc += k3
                if (M2) q_sts(xsb + n * ROW_BYTES, MSED_ADD(x12[n], rhs2[n]));                       // :152  acc = 0.5*k1 + k2
                if (M1) x12[n] = MSED_MUL(0.5, rhs1[n]);                                             // :147  acc = 0.5*k1
            } else {
                if (M3) q_sts(xsa + n * ROW_BYTES, fma(3.0, rhs3[n], xb[n]));                        // :178  Q += 3*k3
                if (M2) {                                                                            // :174  P = k1-k2 ; Q = k1+3*k2
                    q_sts(xsp + n * ROW_BYTES, MSED_SUB(x12[n], rhs2[n]));
                    q_sts(xsb + n * ROW_BYTES, fma(3.0, rhs2[n], x12[n]));
                }
                if (M1) x12[n] = rhs1[n];
            }
        }
        kmod = (kmod == RKQ_RING - 1) ? 0 : kmod + 1;
        kmod3 = (kmod3 == 2) ? 0 : kmod3 + 1;
    };

    {
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using I2 = std::integral_constant<int, 2>;
        using I3 = std::integral_constant<int, 3>;
        iteration(I1{}, I0{}, I0{}, I0{}, 0);
        iteration(I2{}, I1{}, I0{}, I0{}, 1);
        iteration(I2{}, I2{}, I1{}, I0{}, 2);
        iteration(I2{}, I2{}, I2{}, I1{}, 3);
#ifdef MSED_RKQUAD_UNROLL
        MSED_UNROLL_PRAGMA(MSED_RKQUAD_UNROLL)
#else
#pragma unroll 1
#endif
        for (int k = 4; k < K - 1; ++k) iteration(I2{}, I2{}, I2{}, I2{}, k);   // steady state
        iteration(I3{}, I2{}, I2{}, I2{}, K - 1);
        iteration(I0{}, I3{}, I2{}, I2{}, K);
        iteration(I0{}, I0{}, I3{}, I2{}, K + 1);
        iteration(I0{}, I0{}, I0{}, I3{}, K + 2);
    }
    cp_async_wait<0>();

    if (nanf) atomicOr(&p.ctl->flags[1], 1);
}
