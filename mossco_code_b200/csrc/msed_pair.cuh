// msed_pair.cuh -- included inside namespace msed after msed_column.cuh.
//
// Two consecutive accepted sub-steps of the Euler / adaptive-Euler integrator in one pass over HBM.
//
// The single-step kernel is HBM-bound (state read once, written once per attempt) with the fp64 pipe
// only half busy.  Columns are independent and the stencil is one layer wide, so the second sub-step can
// chase the first one down the column with a lag of one layer: when stage A has produced layer k,
// stage B has everything it needs for layer k-1.  The intermediate state never leaves the SM (a
// thread-private 4-layer window in shared memory), so a pair costs one read and one write of the
// state instead of two.
//
// What the two stages are is decided by the host's plan of the reference's attempt sequence
// (solver_library.F90:104-140; KParams::pair_kind, msed.cu run_steps):
//   PAIR_FULL   two whole ode_solver calls, each accepted at dt on its first attempt, with the component's
//               check_NaN / minimum clip after each (fabm_sediment_component.F90:1718-1732);
//   PAIR_FIRST  the first two sub-steps of a call that runs at dt_acc = dt/4^depth: stage A evaluates the RHS
//               of the attempts at dt, dt/4, .. that are rejected (:126-128) -- the state is unchanged by a
//               rejection, so it is the same RHS -- raises "rejection seen" for each of them and advances by
//               dt_acc; nothing is clipped inside a call;
//   PAIR_MID    two inner sub-steps;
//   PAIR_LAST   the last two sub-steps of a call; check_NaN / clip after stage B.
// Every stage raises its own violation flag at dt_acc.  plan_controller_kernel commits the launch (or the
// whole group of launches it belongs to) only if the flags say the reference would have taken exactly the
// planned decisions; otherwise nothing is committed (the input buffer is untouched) and the host redoes the
// same attempts one by one -- a fused launch is speculation with a free rollback.  The arithmetic is the
// shared inline code of msed_column.cuh: a committed pair is bit-identical to the single attempts.
//
// Restricted to the hot configuration: bcup_particulate = 1 (no distributed POM flux cascade),
// bioturbation_profile != 3, closed-form porosity (KParams::por_mode 1 or 2).

#ifndef MSED_PAIR_UNROLL
#define MSED_PAIR_UNROLL 2
#endif
#define MSED_PAIR_UNROLL_PRAGMA MSED_UNROLL_PRAGMA(MSED_PAIR_UNROLL)
// (the lambdas of pair_kernel carry always_inline: with six instantiations of the column walk the inliner's budget
//  runs out and the stages would become real calls with their arrays in local memory)
constexpr int PAIR_WIN = 4;  // c1 window slots (3 live layers: j, j+1 and the one being written)
constexpr uint32_t PAIR_STAGE_BYTES = NV * ROW_BYTES;                    // input ring: 8 rows per layer
constexpr uint32_t PAIR_RING_BYTES = RING_STAGES * PAIR_STAGE_BYTES;
constexpr uint32_t PAIR_WIN_BYTES = PAIR_WIN * NV * ROW_BYTES;
constexpr int PAIR_COLROWS = NV + 2;  // per-column scalars: temperature, surface porosity, one upper-boundary value per variable
static_assert(PAIR_COLROWS <= 2 * NV, "the column scalars borrow window slots 2 and 3");
#ifdef MSED_PAIR_PAD_SMEM   // experiment: the same kernel with a smaller L1 (profiles/r02_summary.md)
constexpr size_t PAIR_SMEM_BYTES = (size_t)PAIR_RING_BYTES + PAIR_WIN_BYTES + MSED_PAIR_PAD_SMEM;
#else
constexpr size_t PAIR_SMEM_BYTES = (size_t)PAIR_RING_BYTES + PAIR_WIN_BYTES;
#endif
#ifndef MSED_PAIR_MIN_BLOCKS
#define MSED_PAIR_MIN_BLOCKS (384 / MSED_COL_BLOCK)   // 12 warps per SM: 168 registers per thread
#endif
constexpr int PAIR_MIN_BLOCKS = MSED_PAIR_MIN_BLOCKS;

__device__ __forceinline__ void sts64(uint32_t addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// ---- bulk asynchronous copies (TMA engine, no tensor map: the rows of the state are contiguous) -----------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "MSED_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra MSED_WAIT_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// one box of a 3-D tensor, global -> shared, completion counted in bytes on an mbarrier (UTMALDG)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// How the input state reaches shared memory:
// what a stage does where the component checks for NaN and clips to the minimum
enum PairClip : int { CLIP_NONE = 0, CLIP_NOW, CLIP_DETECT };

enum PairFeed : int {
    FEED_CPASYNC = 0,  // every thread copies its own column, 8 bytes per row (LDGSTS through the L1)
    FEED_COLMAP,       // the same over the wet-column list of a tile with land (columns of a warp are not contiguous)
    FEED_BULK          // TMA: a warp's 32 columns x 8 variables of one layer are one box of the state tensor
                       // [nvar][K][ld]; one cp.async.bulk.tensor per layer and warp, issued by lane 0, completion on a
                       // warp-private mbarrier per ring slot; the data never passes through the L1
};
// FEED_BULK keeps the ring per warp: [warp][slot][variable][32 columns]
constexpr uint32_t BULK_ROW_BYTES = 32 * 8;
constexpr uint32_t BULK_STAGE_BYTES = NV * BULK_ROW_BYTES;
static_assert(BULK_STAGE_BYTES * (COL_BLOCK / 32) == PAIR_STAGE_BYTES, "both ring layouts fill the same shared memory");

template <int MODEL, bool ADAPTIVE, bool DENIT, int FEED = FEED_CPASYNC, bool OVR = false>
__global__ void __launch_bounds__(COL_BLOCK, PAIR_MIN_BLOCKS)
pair_kernel(const __grid_constant__ KParams p)
{
    constexpr bool COLMAP = FEED == FEED_COLMAP;
    constexpr bool BULK = FEED == FEED_BULK;
    extern __shared__ __align__(16) double ring[];
    __shared__ __align__(8) unsigned long long bars[(COL_BLOCK / 32) * RING_STAGES];
    const Ctl *ctl = p.ctl;
    // the plan this launch belongs to was made for one definite control state: after a failed group
    // (pairs_disabled) or anything else unforeseen the launch does nothing
    if (ctl->stop || ctl->pairs_disabled || ctl->steps_done != p.gate_steps) return;
    const int cur = ctl->cur;
    const bool do_clip = ctl->do_clip != 0;
    const double dt = p.dt_acc;
    // a pair whose violation flags are already up cannot be committed: later CTAs skip their work
    const volatile int *flags = ctl->flags;
    {
        bool quit = ADAPTIVE && dt > ctl->dt_min && (flags[0] | flags[2]);
        if (FEED == FEED_BULK) quit = __any_sync(0xffffffffu, quit);   // the warp is fed as a team: it leaves as one
        if (quit) return;
    }

    // On a tile with land the launch runs over the list of wet columns, so that every lane of a warp has
    // a column to integrate (a land lane idles for the whole walk down its neighbours' columns, and a CTA
    // with one wet warp holds a full CTA's registers and shared memory).  Wet neighbours stay neighbours:
    // accesses remain coalesced except where a run of land is skipped.
    const int cta_col0 = p.col0 + blockIdx.x * COL_BLOCK;
    const int lane = threadIdx.x & 31;
    const int warp_col0 = cta_col0 + (int)(threadIdx.x & ~31u);
    int t = cta_col0 + threadIdx.x;
    if (BULK) {
        // the warp is fed as a team: it stays whole, and a lane past the ragged end of the range walks the last
        // column once more (same values to the same addresses).  The host picks this feed only for a tile
        // without land.
        if (warp_col0 >= p.col_end) return;
        t = min(t, p.col_end - 1);
    } else if (t >= p.col_end) {
        return;
    }
    const int col = COLMAP ? p.colmap[t] : t;   // COLMAP: the tile has land (a separate instantiation, so the
                                                // land-free kernel keeps its register allocation)
    // conc of a land column stays missing_value in both buffers; the wet-column list holds no land column
    // (msed_set_mask builds both), so that feed skips the look-up: one dependent trip to L2 less at the head
    if (!BULK && !COLMAP && p.mask[col] != 0) return;

    const int K = p.K;
    const size_t ld = p.ld;
    const size_t plane = (size_t)K * ld;
    // OVR: one launch of a chunk-major sequence (run_steps): the pairs of a coupling interval follow each
    // other chunk by chunk through explicitly named buffers, the committed state stays untouched
    const double *in = (OVR ? p.in_ovr : p.buf[cur]) + col;
    double *out = (OVR ? p.out_ovr : p.buf[1 - cur]) + col;

    // ---- input ring (cp.async, as in column_kernel) and the c1 window --------------------------
    constexpr uint32_t SROW = BULK ? BULK_ROW_BYTES : ROW_BYTES;              // ring: distance between two variables
    constexpr uint32_t SSTAGE = BULK ? BULK_STAGE_BYTES : PAIR_STAGE_BYTES;   // ... and between two slots
    const uint32_t warp_ring = smem_u32(ring) + (threadIdx.x >> 5) * (RING_STAGES * BULK_STAGE_BYTES);
    const uint32_t sbase = BULK ? warp_ring + (uint32_t)(t - warp_col0) * 8u : smem_u32(ring) + threadIdx.x * 8u;
    const uint32_t wbase = smem_u32(ring) + threadIdx.x * 8u + PAIR_RING_BYTES;   // thread-private in every feed
    // The per-column scalars (temperature, surface porosity, the upper-boundary value of every variable) travel
    // with layer 0 through cp.async as well: as plain loads they were ten dependent trips to HBM at the head
    // of every column (the asm statements of the ring pin their order), a seventh of every warp's life
    // (profiles/r02_summary.md).  They land in slots 2 and 3 of the c1 window, which stage A does not write
    // before layer 2, when both upper boundaries have long been evaluated: no shared memory of their own (a
    // larger carve-out leaves the L1 too small for the cp.async lines in flight, same file).
    const uint32_t cbase = wbase + 2 * (NV * ROW_BYTES);
    // FEED_BULK: the warp's ring-slot barriers and the descriptor of the input buffer
    const uint32_t bar0 = smem_u32(bars) + (threadIdx.x >> 5) * (RING_STAGES * 8u);
    const CUtensorMap *tmap = OVR ? &p.tmap[2] : &p.tmap[cur];
    if (BULK) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < RING_STAGES; ++s) mbar_init(bar0 + s * 8u, 1);
            mbar_init_fence();
        }
        __syncwarp();
    }
    auto fetch_column_rows = [&]() __attribute__((always_inline)) {
        cp_async8(cbase, p.bdys + col);                                   // temp3d(:,:,k) = bdys(:,:,1), driver :602
        if (p.por_mode == 2) cp_async8(cbase + ROW_BYTES, p.por + col);   // porosity(:,:,1), driver :411
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const int bc = (n < NPART) ? p.bcup_part : p.bcup_diss;
            if (bc == 1 || bc == 4) cp_async8(cbase + (2 + n) * ROW_BYTES, p.fluxes + (size_t)n * ld + col);   // :783,:792
            else if (bc == 2) cp_async8(cbase + (2 + n) * ROW_BYTES, p.bdys + (size_t)(n + 1) * ld + col);    // :786
        }
    };
    const double *g_in = in;
    int k_fetch = 0;
    auto fetch_next = [&]() __attribute__((always_inline)) {
        if (BULK) {
            if (k_fetch < K) {
                const uint32_t slot = (uint32_t)(k_fetch & (RING_STAGES - 1));
                __syncwarp();   // every lane has finished with the layer that lived in this slot
                if (lane == 0) {
                    mbar_expect_tx(bar0 + slot * 8u, BULK_STAGE_BYTES);   // the whole box counts, clipped or not
                    tma_load_3d(warp_ring + slot * BULK_STAGE_BYTES, tmap, warp_col0, k_fetch, 0, bar0 + slot * 8u);
                }
            }
            ++k_fetch;
            return;
        }
        if (k_fetch < K) {
            const uint32_t sa = sbase + (uint32_t)(k_fetch & (RING_STAGES - 1)) * PAIR_STAGE_BYTES;  // (not BULK)
            const double *g = g_in;
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                cp_async8(sa + n * ROW_BYTES, g);
                g += plane;
            }
            g_in += ld;
        }
        ++k_fetch;
        cp_async_commit();
    };
    double por_surf = 1.0, cpart, cdiss, fT, temp;
    double FA[NV], FB[NV];
    int viol1 = 0, viol2 = 0;  // sign bit = some relative change fell below relative_change_min (violates_acc)
    bool nan1 = false, nan2 = false;
    int neg = 0;               // sign bit = some new value was negative where the component clips (lazy clip, below)
    double *g_out = out;
    double fT_diag = 1.0;
    auto por_at = [&](int kk) __attribute__((always_inline)) -> double { return __dmul_rn(por_surf, p.portab[kk]); };

    // upper boundary of one step: F[n] = Flux(1) (diff3d :782-803), c0(n) = state of layer 1
    auto top_boundary = [&](auto c0, double por0, double (&F)[NV], bool write_fluxes) __attribute__((always_inline)) {
        double Dp, Dd;
        top_coeffs(cpart, cdiss, por0, p.bf[0], Dp, Dd);
        const double rdz0 = 1.0 / p.dz[0];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const bool part = n < NPART;
            const int bc = part ? p.bcup_part : p.bcup_diss;
            double f = 0.0;
            if (bc == 1 || bc == 4) {
                f = lds64(cbase + (2 + n) * ROW_BYTES);
            } else if (bc == 2) {
                const double Cup = lds64(cbase + (2 + n) * ROW_BYTES);
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

    // layer kk has landed (kk < K)
    auto wait_layer = [&](int kk) __attribute__((always_inline)) {
        if (BULK) mbar_wait(bar0 + (uint32_t)(kk & (RING_STAGES - 1)) * 8u, (uint32_t)(kk / RING_STAGES) & 1u);
        else cp_async_wait<RING_STAGES - 2>();
    };
    // (re)start the walk down the column: the input stream from layer 0, the column's scalars and constants,
    // the upper boundary of step 1
    auto start_column = [&]() __attribute__((always_inline)) {
        if (!BULK) cp_async_wait<0>();   // (a restart: nothing of the first walk is still in flight)
        fetch_column_rows();
        g_in = in;
        k_fetch = 0;
        g_out = out;
#pragma unroll
        for (int s = 0; s < RING_STAGES - 1; ++s) fetch_next();
        if (BULK) {
            cp_async_commit();  // the column's scalars
            cp_async_wait<0>();
        }
        wait_layer(0);  // ... and the column's scalars
        if (p.por_mode == 2) por_surf = lds64(cbase + ROW_BYTES);
        temp = lds64(cbase);
        column_constants<MODEL, false>(p, temp, cpart, cdiss, fT);
        // Step 1 reads the particulate input fluxes BEFORE step 2 may overwrite the dissolved entries of
        // the same array; with bcup_dissolved = 1 the dissolved input fluxes are read here as well and
        // rewritten unchanged by step 2 (:692).
        top_boundary([&](int n) { return lds64(sbase + n * SROW); }, por_at(0), FA, false);
        // The last pair of a call leaves the denitrification diagnostic of its second step behind: it
        // describes the state of the last get_rhs call, which here never reaches HBM.
        fT_diag = fT;
        if (MODEL != MSED_MODEL_OMEXDIA_P && DENIT) fT_diag = exp(-p.om.E_a * (1.0 / (temp + 273.15) - 1.0 / 288.15));
    };

    // one step of one layer: finishes layer kk given its state cc, the state cn of the layer below and
    // the flux F through its upper interface; the new state goes to `sink`.  HAS_NEXT and CLIP are
    // compile-time so the steady-state loop body is branch-free and both steps can be interleaved.
    // state-independent coefficients of layer kk and of its lower interface: both steps of the pair
    // need the same ones (step 2 one iteration later), so they are computed once
    struct LayerCoef { double porc, porn, mDp, mDd, rpd; };
    auto make_coef = [&](auto has_next_tag, int kk) __attribute__((always_inline)) -> LayerCoef {
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

    // planned rejections (PAIR_FIRST, stage A): bit l of viol_up = some relative change at dt_acc*4^(l+1) fell
    // below relative_change_min.  Tested exactly as the single attempt tests it (violates(): a false alarm
    // here would commit a sub-cycle the reference never took).
    int viol_up = 0;
    static_assert(MAX_PLAN_DEPTH == 2, "dt_up lists the step sizes of the planned rejections");
    const double dt_up[MAX_PLAN_DEPTH] = {dt * 4.0, dt * 16.0};   // exact: dt_acc = dt_up * 0.25 (:127)
    const int depth = p.depth;

    auto step_layer = [&](auto has_next_tag, auto clip_tag, auto up_tag, const LayerCoef &lc, const double (&cc)[NV],
                          auto cn, double (&F)[NV], int &viol, bool &nanf, auto sink, auto denit) __attribute__((always_inline)) {
        constexpr bool HAS_NEXT = decltype(has_next_tag)::value;
        constexpr int CLIPM = decltype(clip_tag)::value;   // CLIP_NONE / CLIP_NOW / CLIP_DETECT
        constexpr bool UP = decltype(up_tag)::value;
        double Fn[NV];
        if (HAS_NEXT) {
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
            omexdia_rates(p.om, cc, fT, r, denit);
        } else {
            // what field_kernel reports for this model
            if (!std::is_same<decltype(denit), std::nullptr_t>::value) omexdia_rates(p.om, cc, fT_diag, r, denit);
#pragma unroll
            for (int n = 0; n < NV; ++n) r[n] = 0.0;
        }
        double raw[NV];  // new state before the clip (check_NaN looks at it first, component :1718)
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const double rhs = layer_rhs(F[n], Fn[n], lc.rpd, r[n]);
            F[n] = Fn[n];
            const double c0 = cc[n];
            double newc = euler_update(dt, rhs, c0);
            if (ADAPTIVE) violates_acc(viol, p.fac, c0, newc);
            if (ADAPTIVE && UP) {
#pragma unroll
                for (int l = 0; l < MAX_PLAN_DEPTH; ++l)
                    if (l < depth && violates(p.fac, c0, euler_update(dt_up[l], rhs, c0))) viol_up |= 1 << l;
            }
            raw[n] = newc;
            if (CLIPM != CLIP_NONE) {
                if (n & 1) nanf |= either_nan(raw[n - 1], raw[n]);
                if (CLIPM == CLIP_NOW) newc = clip_min(newc, p.om.minimum[n]);
                else neg |= __double2hiint(newc);   // (merged by the compiler into one LOP3 per two values)
            }
            sink(n, newc);
        }
    };

    LayerCoef coef_prev;  // coefficients of the layer step 2 is about to process (made by step 1)
    // step 1, layer k: state from the ring, result into the c1 window; returns the layer coefficients
    auto stage_a = [&](auto has_next_tag, auto clip_tag, auto up_tag, int k) __attribute__((always_inline)) -> LayerCoef {
        fetch_next();
        if (!BULK || decltype(has_next_tag)::value) wait_layer(k + 1);
        const uint32_t sc = sbase + (uint32_t)(k & (RING_STAGES - 1)) * SSTAGE;
        const uint32_t sn = sbase + (uint32_t)((k + 1) & (RING_STAGES - 1)) * SSTAGE;
        const uint32_t wk = wbase + (uint32_t)(k & (PAIR_WIN - 1)) * (NV * ROW_BYTES);
        const LayerCoef lc = make_coef(has_next_tag, k);
        double cc[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) cc[n] = lds64(sc + n * SROW);
        step_layer(has_next_tag, clip_tag, up_tag, lc, cc, [&](int n) { return lds64(sn + n * SROW); }, FA,
                   viol1, nan1, [&](int n, double v) { sts64(wk + n * ROW_BYTES, v); }, nullptr);
        return lc;
    };
    // step 2, layer j: state from the c1 window, result to HBM
#ifdef MSED_PAIR_PRELOAD_B
    // (experiment) In the steady state the layer stage B evaluates was written by stage A one iteration ago: read
    // ahead of stage A's stores of this iteration, it lets B's reaction rates run beside A's arithmetic -- the
    // ld.shared / st.shared statements keep their order, so a load behind A's stores waits for all of A.
    double ccB[NV];
    auto preload_b = [&](int j) __attribute__((always_inline)) {
        const uint32_t wj = wbase + (uint32_t)(j & (PAIR_WIN - 1)) * (NV * ROW_BYTES);
#pragma unroll
        for (int n = 0; n < NV; ++n) ccB[n] = lds64(wj + n * ROW_BYTES);
    };
#endif
    auto stage_b = [&](auto has_next_tag, auto clip_tag, int j, const LayerCoef &lc, auto preloaded_tag) __attribute__((always_inline)) {
        const uint32_t wj = wbase + (uint32_t)(j & (PAIR_WIN - 1)) * (NV * ROW_BYTES);
        const uint32_t wn = wbase + (uint32_t)((j + 1) & (PAIR_WIN - 1)) * (NV * ROW_BYTES);
        double cc[NV];
#ifdef MSED_PAIR_PRELOAD_B
        if (decltype(preloaded_tag)::value) {
#pragma unroll
            for (int n = 0; n < NV; ++n) cc[n] = ccB[n];
        } else
#endif
        {
#pragma unroll
            for (int n = 0; n < NV; ++n) cc[n] = lds64(wj + n * ROW_BYTES);
        }
        double *go = g_out;
        if (DENIT) {
            double dn = 0.0;
            step_layer(has_next_tag, clip_tag, std::false_type{}, lc, cc, [&](int n) { return lds64(wn + n * ROW_BYTES); },
                       FB, viol2, nan2, [&](int n, double v) { go[(size_t)n * plane] = v; }, &dn);
            p.denit_out[(size_t)j * ld + col] = dn;
        } else {
            step_layer(has_next_tag, clip_tag, std::false_type{}, lc, cc, [&](int n) { return lds64(wn + n * ROW_BYTES); },
                       FB, viol2, nan2, [&](int n, double v) { go[(size_t)n * plane] = v; }, nullptr);
        }
        g_out += ld;
    };

    // clip_a / clip_b: check_NaN + minimum clip after stage A / stage B (the stage ends an ode_solver call and
    // the component wrapper is on); up_tag: stage A also tests the planned rejections
    auto sweep = [&](auto clip_a, auto clip_b, auto up_tag) __attribute__((always_inline)) {
        using Y = std::true_type;
        using N = std::false_type;
        start_column();
        if (K == 1) {  // degenerate column: both stages see a closed bottom right away
            coef_prev = stage_a(N{}, clip_a, up_tag, 0);
            top_boundary([&](int n) { return lds64(wbase + n * ROW_BYTES); }, por_at(0), FB, true);
            stage_b(N{}, clip_b, 0, coef_prev, N{});
            return;
        }
        coef_prev = stage_a(Y{}, clip_a, up_tag, 0);
        top_boundary([&](int n) { return lds64(wbase + n * ROW_BYTES); }, por_at(0), FB, true);
        MSED_PAIR_UNROLL_PRAGMA
        for (int k = 1; k < K - 1; ++k) {  // steady state: stage A on layer k, stage B on layer k-1
#ifdef MSED_PAIR_PRELOAD_B
            preload_b(k - 1);
#endif
            const LayerCoef lc = stage_a(Y{}, clip_a, up_tag, k);
            stage_b(Y{}, clip_b, k - 1, coef_prev, Y{});
            coef_prev = lc;
        }
        const LayerCoef last = stage_a(N{}, clip_a, up_tag, K - 1);
        stage_b(Y{}, clip_b, K - 2, coef_prev, N{});
        stage_b(N{}, clip_b, K - 1, last, N{});
    };
    {
        using Y = std::true_type;
        using N = std::false_type;
        const int kind = ADAPTIVE ? p.pair_kind : (int)PAIR_FULL;
        // one call site per instantiation (the lambda body is inlined wherever it is called)
        const int sel = (kind == PAIR_FULL && do_clip) ? 0 : (kind == PAIR_FIRST) ? 1 : (kind == PAIR_LAST && do_clip) ? 2 : 3;
        using C0 = std::integral_constant<int, CLIP_NONE>;
        using C1 = std::integral_constant<int, CLIP_NOW>;
        using CD = std::integral_constant<int, CLIP_DETECT>;
        // Lazy clip.  The component's clip to the minimum (fabm_sediment_component.F90:1728-1730) is a safety net
        // that almost never fires, but as four integer instructions per value it is an eighth of the loop.  With
        // zero minima (the FABM default) "some value would be clipped" is "some value has its sign bit set": the
        // column is walked without the clip, the sign bits are collected, and a thread that did meet a negative
        // value (or -0.0) walks its column again with the clip in place -- the input buffer is untouched, every
        // output is simply written again.  Same bits either way.
        // (the reaction-free test model keeps the plain clip: two walks fewer to compile)
        constexpr bool LAZY = MODEL == MSED_MODEL_OMEXDIA_P && !BULK;
        const bool lazy = LAZY && p.min_zero != 0;
        switch (sel) {
        case 0:
            if constexpr (LAZY) {
                if (lazy) {
                    sweep(CD{}, CD{}, N{});
                    if (neg >= 0) break;
                }
            }
            sweep(C1{}, C1{}, N{});   // no lazy clip, or the thread met a value the clip changes
            break;
        case 1: sweep(C0{}, C0{}, Y{}); break;
        case 2:
            if constexpr (LAZY) {
                if (lazy) {
                    sweep(C0{}, CD{}, N{});
                    if (neg >= 0) break;
                }
            }
            sweep(C0{}, C1{}, N{});
            break;
        default: sweep(C0{}, C0{}, N{}); break;
        }
    }
    if (!BULK) cp_async_wait<0>();

    int *wf = p.ctl->flags;
    if (ADAPTIVE && viol1 < 0) atomicOr(&wf[0], 1);
    if (nan1) atomicOr(&wf[1], 1);
    if (ADAPTIVE && viol2 < 0) atomicOr(&wf[2], 1);
    if (nan2) atomicOr(&wf[3], 1);
    if (ADAPTIVE && viol_up) {
#pragma unroll
        for (int l = 0; l < MAX_PLAN_DEPTH; ++l)
            if (viol_up & (1 << l)) atomicOr(&wf[FLAG_UP0 + p.up_slot + l], 1);
    }
}

