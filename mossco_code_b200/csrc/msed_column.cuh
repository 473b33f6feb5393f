// msed_column.cuh -- included inside namespace msed by msed_kernels.cuh.
//
// The fused column kernel.  One thread owns one sediment column (lanes of a warp own adjacent
// columns -> every global access of the [nvar][knum][ld] state is a coalesced 256-byte warp access)
// and walks it top -> bottom.  Global reads never stall the math: each thread streams its column
// through a private STAGES-deep ring in shared memory with cp.async (LDGSTS), STAGES-1 layers ahead
// of the layer it is computing, so the loads in flight per SM are set by the ring depth, not by the
// register file.  Per layer the thread reads layer k and k+1 back from its ring slots, evaluates
//   - the diff3d interface flux (fabm_sediment_driver.F90:776-778) with the layer-k+1 diffusivities,
//   - the omexdia_p reaction rates of layer k,
//   - dC (:819), the integrator update (solver_library.F90:102/:111/:147-182), the relative-change
//     test (:121), check_NaN (component :2392) and the minimum clip (component :1728),
// and stores layer k.  Per ode_solver attempt the state is read once and written once.

// ---- small device helpers ---------------------------------------------------------------------
__device__ __forceinline__ double ld_ro(const double *p) { return __ldg(p); }

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// 1/x to within 2 ulp: MUFU.RCP64H seed (measured 2^-19.9) + one cubic step (3 DFMA); two Newton
// steps (4 DFMA) give 1.7 ulp -- tools/rcp_accuracy.cu measures both.  No IEEE slow path: every
// denominator on this path is a positive, normal number (half-saturation sums, porosity*dz).
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
#ifndef MSED_RCP_NEWTON2
    y = fma(y, fma(e, e, e), y);  // one cubic step: y*(1 + e + e^2)
#else
    y = fma(y, e, y);             // two Newton steps
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
#endif
    return y;
}
// true if a or b is NaN: one DSETP for two values
__device__ __forceinline__ bool either_nan(double a, double b)
{
    int r;
    asm("{\n\t.reg .pred p;\n\tsetp.nan.f64 p, %1, %2;\n\tselp.s32 %0, 1, 0, p;\n\t}" : "=r"(r) : "d"(a), "d"(b));
    return r != 0;
}

// Every product, sum and fused multiply-add below is spelled out (DMUL/DADD/DFMA intrinsics), so the
// compiler has no contraction freedom left: the same source gives the same bits in every kernel that
// inlines it (single step, fused pair, RK stages, the diagnostics kernel), whatever else those kernels
// use the intermediate values for.  Bit-identity between the fused and unfused paths rests on this.
#define MSED_MUL(a, b) __dmul_rn((a), (b))
#define MSED_ADD(a, b) __dadd_rn((a), (b))
#define MSED_SUB(a, b) __dsub_rn((a), (b))

// two reciprocals from one refinement: 1/a = b/(a*b), 1/b = a/(a*b)
__device__ __forceinline__ void fast_rcp2(double a, double b, double &ra, double &rb)
{
    const double y = fast_rcp(MSED_MUL(a, b));
    ra = MSED_MUL(y, b);
    rb = MSED_MUL(y, a);
}

// hzg_omexdia_p local rates for one cell: SURVEY.md Appendix B (frozen project spec; the FABM
// source is not part of the reference tree).  fT is the per-column Arrhenius factor; every a/b of
// the spec is evaluated as a*fast_rcp(b).
__device__ __forceinline__ void omexdia_rates(const OmexDev &m, const double (&c)[NV], double fT,
                                              double (&r)[NV], double *denit)
{
    const double ldetC = c[0], sdetC = c[1], detP = c[2], po4 = c[3];
    const double no3 = c[4], nh3 = c[5], oxy = c[6], odu = c[7];
    const double relaxO2 = 0.04;

#ifndef MSED_RATES_V1
    // The three limitation terms over a common denominator (algebraically the spec's expressions):
    //   Oxicminlim = oxy/d1,  Denitrilim = (1 - oxy/d2) no3/d3 = kinO2denit no3/(d2 d3),
    //   Anoxiclim = (1 - oxy/d4)(1 - no3/d5) = kinO2anox kinNO3anox/(d4 d5),  Rescale = 1/(their sum)
    // so that lim*Rescale = t_i/(t1+t2+t3) with t1 = oxy d2d3 d4d5, t2 = kinO2denit no3 d1 d4d5,
    // t3 = kinO2anox kinNO3anox d1 d2d3: ONE reciprocal where the term-by-term form needs five, and the
    // dependent chain from the state to the rates is ten operations shorter (the fused kernels are bound by
    // the fp64 pipe and by dependent-issue latency).  Differences to the term-by-term rounding are O(1e-16).
    const double d1 = fma(relaxO2, MSED_ADD(nh3, odu), MSED_ADD(oxy, m.ksO2oxic));
    const double d2 = MSED_ADD(oxy, m.kinO2denit), d3 = MSED_ADD(no3, m.ksNO3denit);
    const double d4 = MSED_ADD(oxy, m.kinO2anox), d5 = MSED_ADD(no3, m.kinNO3anox);
    double r7, r8;
    fast_rcp2(fma(relaxO2, MSED_ADD(ldetC, odu), MSED_ADD(oxy, m.ksO2nitri)),
              fma(relaxO2, MSED_ADD(nh3, ldetC), MSED_ADD(oxy, m.ksO2oduox)), r7, r8);
    const double P23 = MSED_MUL(d2, d3), P45 = MSED_MUL(d4, d5);
    const double t1 = MSED_MUL(MSED_MUL(oxy, P23), P45);
    const double t2 = MSED_MUL(MSED_MUL(MSED_MUL(m.kinO2denit, no3), d1), P45);
    const double t3 = MSED_MUL(MSED_MUL(MSED_MUL(m.kinO2anox, m.kinNO3anox), d1), P23);
    const double rN = fast_rcp(MSED_ADD(MSED_ADD(t1, t2), t3));
    const double Oxicminlim = MSED_MUL(oxy, fast_rcp(d1));

    const double CprodL = MSED_MUL(m.rLabile, ldetC);
    const double CprodS = MSED_MUL(m.rSemilabile, sdetC);
    const double Csum = MSED_ADD(CprodL, CprodS);
    const double Cprod = (Csum > m.CprodMax) ? m.CprodMax : Csum;
    const double Nprod = fma(CprodS, m.NCrSdet, MSED_MUL(CprodL, m.NCrLdet));

    const double radsP = MSED_MUL(MSED_MUL(m.PAds_rS, po4), (odu > m.PAdsODU) ? odu : m.PAdsODU);
    const double rP = MSED_MUL(m.rLabile, MSED_SUB(1.0, Oxicminlim));

    const double CN = MSED_MUL(Cprod, rN);                  // Cprod*Rescale/(d1 d2d3 d4d5)
    const double OxicMin = MSED_MUL(CN, t1);
    const double Denitrific = MSED_MUL(CN, t2);
    const double AnoxicMin = MSED_MUL(CN, t3);
#else
    const double r1 = fast_rcp(fma(relaxO2, MSED_ADD(nh3, odu), MSED_ADD(oxy, m.ksO2oxic)));
    double r2, r3, r4, r5, r7, r8;  // paired: every denominator is a positive half-saturation sum
    fast_rcp2(MSED_ADD(oxy, m.kinO2denit), MSED_ADD(oxy, m.kinO2anox), r2, r4);
    fast_rcp2(MSED_ADD(no3, m.ksNO3denit), MSED_ADD(no3, m.kinNO3anox), r3, r5);
    fast_rcp2(fma(relaxO2, MSED_ADD(ldetC, odu), MSED_ADD(oxy, m.ksO2nitri)),
              fma(relaxO2, MSED_ADD(nh3, ldetC), MSED_ADD(oxy, m.ksO2oduox)), r7, r8);

    const double Oxicminlim = MSED_MUL(oxy, r1);
    const double Denitrilim = MSED_MUL(MSED_MUL(fma(-oxy, r2, 1.0), no3), r3);
    const double Anoxiclim = MSED_MUL(fma(-oxy, r4, 1.0), fma(-no3, r5, 1.0));
    const double Rescale = fast_rcp(MSED_ADD(MSED_ADD(Oxicminlim, Denitrilim), Anoxiclim));

    const double CprodL = MSED_MUL(m.rLabile, ldetC);
    const double CprodS = MSED_MUL(m.rSemilabile, sdetC);
    const double Csum = MSED_ADD(CprodL, CprodS);
    const double Cprod = (Csum > m.CprodMax) ? m.CprodMax : Csum;
    const double Nprod = fma(CprodS, m.NCrSdet, MSED_MUL(CprodL, m.NCrLdet));

    const double radsP = MSED_MUL(MSED_MUL(m.PAds_rS, po4), (odu > m.PAdsODU) ? odu : m.PAdsODU);
    const double rP = MSED_MUL(m.rLabile, MSED_SUB(1.0, Oxicminlim));

    const double CR = MSED_MUL(Cprod, Rescale);
    const double Denitrific = MSED_MUL(CR, Denitrilim);
    const double OxicMin = MSED_MUL(CR, Oxicminlim);
    const double AnoxicMin = MSED_MUL(CR, Anoxiclim);
#endif

    const double Nitri = MSED_MUL(MSED_MUL(MSED_MUL(MSED_MUL(fT, m.rnit), nh3), oxy), r7);
    const double OduOx = MSED_MUL(MSED_MUL(MSED_MUL(MSED_MUL(fT, m.rODUox), odu), oxy), r8);

    r[0] = MSED_MUL(-fT, CprodL);
    r[1] = MSED_MUL(-fT, CprodS);
    r[2] = MSED_MUL(fT, fma(-rP, detP, radsP));                   // fT*(radsP - Pprod), Pprod = rP*detP
    r[3] = -r[2];  // = fT * (Pprod - radsP) exactly
    r[4] = fma(-0.8, Denitrific, Nitri);
    r[5] = MSED_MUL(MSED_SUB(Nprod, Nitri), m.rNH3Ads);
    r[6] = MSED_SUB(fma(-2.0, Nitri, -OduOx), OxicMin);           // -OxicMin - 2 Nitri - OduOx
    r[7] = MSED_SUB(AnoxicMin, OduOx);                            // AnoxicMin - OduOx
    if (denit) *denit = MSED_MUL(0.8, Denitrific);
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

// ---- the transport arithmetic, shared by column_kernel and pair_kernel (msed_pair.cuh) so that a
// fused pair of steps is bit-identical to two single steps ---------------------------------------
// diffusivities of the interface between two layers, pre-multiplied by -1/dzc
// (driver :652-653 bioturbation part, :682-683 molecular part, diff3d :777)
__device__ __forceinline__ void interface_coeffs(double cpart, double cdiss, double porc, double porn,
                                                 double bfk, double rdzc, double &mDp, double &mDd)
{
    const double intf = MSED_MUL(0.5, MSED_ADD(porc, porn));  // intf_porosity, driver :435
    const double Dp = MSED_MUL(MSED_MUL(cpart, MSED_SUB(1.0, intf)), bfk);
    const double Dd = fma(cdiss, intf, Dp);
    mDp = MSED_MUL(-Dp, rdzc);
    mDd = MSED_MUL(-Dd, rdzc);
}
__device__ __forceinline__ double flux_particulate(double mDp, double cn, double porn, double cc, double porc)
{
    return MSED_MUL(mDp, fma(cn, porn, -MSED_MUL(cc, porc)));  // C = conc*porosity, driver :663
}
__device__ __forceinline__ double flux_dissolved(double mDd, double cn, double cc)
{
    return MSED_MUL(mDd, MSED_SUB(cn, cc));
}
// upper-boundary diffusivities: intf_porosity(:,:,1) = porosity(:,:,1), driver :434
__device__ __forceinline__ void top_coeffs(double cpart, double cdiss, double por0, double bf0, double &Dp,
                                           double &Dd)
{
    Dp = MSED_MUL(MSED_MUL(cpart, MSED_SUB(1.0, por0)), bf0);
    Dd = fma(cdiss, por0, Dp);
}
__device__ __forceinline__ double top_flux_dirichlet(double D, double C1, double Cup, double rdz0)
{
    return MSED_MUL(MSED_MUL(-D, MSED_SUB(C1, Cup)), rdz0);  // diff3d :786
}
// dC (:819) with the particulate rescaling (:677-678) folded: (Flux(k)-Flux(k+1))/(porosity*dz) + rate
__device__ __forceinline__ double layer_rhs(double Fup, double Flow, double rpd, double rate)
{
    return fma(MSED_SUB(Fup, Flow), rpd, rate);  // driver :715
}
__device__ __forceinline__ double euler_update(double dt, double rhs, double c0) { return fma(dt, rhs, c0); }
__device__ __forceinline__ bool violates(double fac, double c0, double newc)  // solver_library.F90:121
{
    return fma(-fac, c0, newc) < 0.0;
}
// The same test for the speculative kernels (pairs, chains), accumulated on the integer pipe: a negative
// value has its sign bit set, so the high words are OR-ed and the sign of the result is looked at once.
// -0.0 and negative NaNs set the bit too; in a speculative launch a false alarm only turns a commit into
// a redo by the single-step kernel, which applies violates() itself -- results cannot change, and the
// fp64 pipe, which bounds those kernels, is spared a DSETP per variable and step.
__device__ __forceinline__ void violates_acc(int &acc, double fac, double c0, double newc)
{
    acc |= __double2hiint(fma(-fac, c0, newc));
}
// The component's clip, conc = max(conc, minimum) (fabm_sediment_component.F90:1728-1730), on the integer pipe: for a
// minimum >= +0 "v < mn" is the signed comparison of the two bit patterns (non-negative doubles are ordered like
// their bits, every negative double has the sign bit set and compares below) -- two ISETP instead of a DSETP on
// the fp64 pipe, which bounds the fused kernels.  msed_create rejects a negative minimum and turns -0 into +0.
// Where it differs from the floating-point test nothing observable changes: -0.0 becomes +0.0 under a zero
// minimum, and a NaN with the sign bit set is replaced -- check_NaN looks at the value BEFORE the clip (:1718).
__device__ __forceinline__ double clip_min(double v, double mn)
{
    return __double_as_longlong(v) < __double_as_longlong(mn) ? mn : v;
}
// per-column Arrhenius factors and diffusivity prefactors (driver :648,:652,:682; omexdia_p f_T)
template <int MODEL, bool PROFILE3>
__device__ __forceinline__ void column_constants(const KParams &p, double temp, double &cpart, double &cdiss,
                                                 double &fT)
{
    fT = 1.0;
    if (PROFILE3) {
        cpart = 1.0 / 86400.0 / 10000.0;  // f_T = 1, bioturbation = 1, driver :622-623
    } else {
        const double f_T = exp(-4500.0 * (1.0 / (temp + 273.0) - (1.0 / 288.0)));  // :648
        cpart = p.bioturbation * f_T / 86400.0 / 10000.0;                          // :652
    }
    cdiss = fma(temp, 0.035, p.diffusivity) / 86400.0 / 10000.0;                   // :682-683
    if (MODEL == MSED_MODEL_OMEXDIA_P) fT = exp(-p.om.E_a * (1.0 / (temp + 273.15) - 1.0 / 288.15));
}

template <int OP> struct OpTraits {
    static constexpr bool stepping = (OP != OP_RHS);
    static constexpr bool reads_base = (OP == OP_RK4_S2 || OP == OP_RK4_S3 || OP == OP_RK4_S4 ||
                                        OP == OP_RK38_S2 || OP == OP_RK38_S3 || OP == OP_RK38_S4);
    static constexpr bool final_stage = (OP == OP_EULER || OP == OP_ADAPTIVE || OP == OP_RK4_S4 ||
                                         OP == OP_RK38_S4);
    // stage evaluates the RHS on the scratch state c1 (buf[1-cur]) rather than on conc
    static constexpr bool in_is_c1 = reads_base;
    static constexpr bool reads_a1 = (OP == OP_RK4_S2 || OP == OP_RK4_S3 || OP == OP_RK4_S4 ||
                                      OP == OP_RK38_S2 || OP == OP_RK38_S3);
    static constexpr bool reads_a2 = (OP == OP_RK38_S3 || OP == OP_RK38_S4);
};

constexpr int ROWS = NV + 1;                 // 8 state rows + porosity per layer
constexpr int RING_STAGES = MSED_RING_STAGES; // layers resident in the shared-memory ring (power of 2)
constexpr uint32_t ROW_BYTES = COL_BLOCK * 8;
constexpr uint32_t STAGE_BYTES = ROWS * ROW_BYTES;
constexpr size_t COLUMN_SMEM_BYTES = (size_t)RING_STAGES * STAGE_BYTES;

// ---------------------------------------------------------------------------------------------
// the column kernel
// ---------------------------------------------------------------------------------------------
// STREAM_POR: porosity is an arbitrary 3-D field streamed through the ring (KParams::por_mode 0);
// otherwise it is por_surf * portab[k] (modes 1 and 2, see below), which costs no HBM traffic.
#define MSED_PRAGMA_STR(x) _Pragma(#x)
#define MSED_UNROLL_PRAGMA(n) MSED_PRAGMA_STR(unroll n)
#ifdef MSED_COL_UNROLL
#define MSED_COL_UNROLL_PRAGMA MSED_UNROLL_PRAGMA(MSED_COL_UNROLL)
#else
#define MSED_COL_UNROLL_PRAGMA
#endif

template <int MODEL, int OP, bool PROFILE3, bool STREAM_POR>
__global__ void __launch_bounds__(COL_BLOCK, COL_MIN_BLOCKS)
column_kernel(const __grid_constant__ KParams p)
{
    using T = OpTraits<OP>;
    extern __shared__ __align__(16) double ring[];
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
            // An attempt whose violation flag is already up will be rejected for certain when
            // dt_red > dt_min (solver_library.F90:126) and its output discarded: CTAs that start
            // after the first violation skip their work, so a rejected attempt costs about one wave.
            if (dt > ctl->dt_min && *(const volatile int *)&ctl->flags[0]) return;
        } else {
            dt = ctl->dt;
        }
    }
    const int col = p.col0 + blockIdx.x * COL_BLOCK + threadIdx.x;
    if (col >= p.col_end) return;

    const int K = p.K;
    const size_t ld = p.ld;
    const size_t plane = (size_t)K * ld;  // distance between two state variables
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

    // ---- thread-private prefetch ring: layer kk lives in slot kk % RING_STAGES --------------------
    const uint32_t sbase = smem_u32(ring) + threadIdx.x * 8u;
    const double *g_in = in, *g_por = por;  // running source pointers of the next layer to fetch
    int k_fetch = 0;
    auto fetch_next = [&]() {
        if (k_fetch < K) {
            const uint32_t sa = sbase + (uint32_t)(k_fetch & (RING_STAGES - 1)) * STAGE_BYTES;
            const double *g = g_in;
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                cp_async8(sa + n * ROW_BYTES, g);
                g += plane;
            }
            if (STREAM_POR) cp_async8(sa + NV * ROW_BYTES, g_por);
            g_in += ld;
            g_por += ld;
        }
        ++k_fetch;
        cp_async_commit();  // committed even when empty so the group accounting stays uniform
    };
#pragma unroll
    for (int s = 0; s < RING_STAGES - 1; ++s) fetch_next();

    // porosity of layer kk: the 3-D field (mode 0, streamed through the ring), the uniform profile
    // of initialize (mode 1, driver :280) or surface value x depth profile (mode 2, driver :411-412) --
    // the last two reproduce the stored field bit for bit and save its 8 B/cell-update of HBM traffic
    // (mode 1 is mode 2 with a surface factor of exactly 1.0: x*1.0 is exact)
    const double por_surf = (!STREAM_POR && p.por_mode == 2) ? ld_ro(por) : 1.0;
    auto por_at = [&](int kk, uint32_t slot) -> double {
        if (STREAM_POR) return lds64(slot + NV * ROW_BYTES);
        return __dmul_rn(por_surf, p.portab[kk]);
    };

    // ---- per-column constants -----------------------------------------------------------
    const double temp = ld_ro(p.bdys + col);  // temp3d(:,:,k) = bdys(:,:,1), driver :602
    double cpart, cdiss, fT;
    column_constants<MODEL, PROFILE3>(p, temp, cpart, cdiss, fT);

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
    cp_async_wait<RING_STAGES - 2>();  // layer 0 has landed
    double F[NV];                      // flux through the upper interface of the current layer
    double rest[NPART];
    bool casc[NPART];
    double cap_prev = 0.0;
    {
        const double por0 = por_at(0, sbase);
        double bf0 = p.bf[0];
        if (PROFILE3) bf0 = bf3_cell(p, 0, wtoc_cell(p, por0, poc[0], poc[plane]), avg_wt);
        double Dp, Dd;
        top_coeffs(cpart, cdiss, por0, bf0, Dp, Dd);
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
                const double c1 = lds64(sbase + n * ROW_BYTES);
                const double C1 = part ? MSED_MUL(c1, por0) : c1;
                f = top_flux_dirichlet(part ? Dp : Dd, C1, Cup, rdz0);
            } else if (bc != 3 && n > 0) {
                // BcUp outside 1..4 (bcup_dissolved_variables = 0): diff3d never assigns Flux(1)
                // (:782-803), so it keeps what the previous variable's call left in get_rhs's
                // intFlux array (:586) -- reproduced here; BcUp = 3 is the explicit zero (:789)
                f = F[n - 1];
            }
            F[n] = f;
            if (!part) p.fluxes[(size_t)n * ld + col] = f;   // fluxes(:,:,n) = intFlux(:,:,1), :692
        }
        if (p.bcup_part == 4) {  // :792-794
            cap_prev = p.pom_flux_rate * (1.0 - por0) * p.dz[0];
#pragma unroll
            for (int n = 0; n < NPART; ++n) {
                rest[n] = F[n] - cap_prev;
                casc[n] = true;
                if (K == 1) F[n] += rest[n];  // k=2 > knum: :800-802
            }
        }
    }

    bool viol = false, nanf = false;
    const bool clip_now = T::final_stage && do_clip && final_sub;
    double *g_out = out;                                   // running pointers of layer k
    const double *g_base = base;
    double *g_a1 = aux1, *g_a2 = aux2;
    double *g_rhs = p.rhs_out + col;

    MSED_COL_UNROLL_PRAGMA
    for (int k = 0; k < K; ++k) {
        const bool has_next = (k + 1 < K);
        fetch_next();                          // layer k+RING_STAGES-1 -> the slot layer k-1 just left
        cp_async_wait<RING_STAGES - 2>();      // layer k+1 has landed
        const uint32_t sc = sbase + (uint32_t)(k & (RING_STAGES - 1)) * STAGE_BYTES;
        const uint32_t sn = sbase + (uint32_t)((k + 1) & (RING_STAGES - 1)) * STAGE_BYTES;

        double cc[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) cc[n] = lds64(sc + n * ROW_BYTES);
        const double porc = por_at(k, sc);

        double basev[NV], a1[NV], a2[NV];
        if (T::reads_base) {
            const double *g = g_base;
#pragma unroll
            for (int n = 0; n < NV; ++n) { basev[n] = *g; g += plane; }
        }
        if (T::reads_a1) {
            const double *g = g_a1;
#pragma unroll
            for (int n = 0; n < NV; ++n) { a1[n] = *g; g += plane; }
        }
        if (T::reads_a2) {
            const double *g = g_a2;
#pragma unroll
            for (int n = 0; n < NV; ++n) { a2[n] = *g; g += plane; }
        }

        // flux through the lower interface (diff3d :776-778; BcDown = 3, :590,:813)
        double Fn[NV];
        if (has_next) {
            const double porn = por_at(k + 1, sn);
            double bfk = p.bf[k + 1];
            if (PROFILE3)
                bfk = bf3_cell(p, k + 1,
                               wtoc_cell(p, porn, poc[(size_t)(k + 1) * ld], poc[plane + (size_t)(k + 1) * ld]),
                               avg_wt);
            double mDp, mDd;
            interface_coeffs(cpart, cdiss, porc, porn, bfk, p.rdzc[k], mDp, mDd);
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const double cn = lds64(sn + n * ROW_BYTES);
                if (n < NPART) Fn[n] = flux_particulate(mDp, cn, porn, cc[n], porc);
                else Fn[n] = flux_dissolved(mDd, cn, cc[n]);
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

        // local reaction rates (fabm_do, driver :700)
        double r[NV];
        if (MODEL == MSED_MODEL_OMEXDIA_P) {
            omexdia_rates(p.om, cc, fT, r, nullptr);
        } else {
#pragma unroll
            for (int n = 0; n < NV; ++n) r[n] = 0.0;
        }

        // dC (:819) with the particulate rescaling (:677-678) folded: both reduce to
        // (Flux(k)-Flux(k+1)) / (porosity*dz)
        const double rpd = fast_rcp(MSED_MUL(porc, p.dz[k]));
        auto finish_layer = [&](auto clip_tag) {
            constexpr bool CLIP = decltype(clip_tag)::value;
            double *go = g_out, *ga1 = g_a1, *ga2 = g_a2, *gr = g_rhs;
            double raw[NV];  // new state before the clip
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const double rhs = layer_rhs(F[n], Fn[n], rpd, r[n]);
                F[n] = Fn[n];
                const double c0 = cc[n];
                double newc = 0.0;
                if (OP == OP_RHS) {
                    *gr = rhs;
                } else if (OP == OP_EULER || OP == OP_ADAPTIVE) {
                    newc = euler_update(dt, rhs, c0);                      // :102,:111
                    if (OP == OP_ADAPTIVE) viol |= violates(p.fac, c0, newc);  // :121
                } else if (OP == OP_RK4_S1) {                               // :147
                    newc = fma(0.5 * dt, rhs, c0);
                    *ga1 = 0.5 * rhs;
                } else if (OP == OP_RK4_S2) {                               // :152
                    newc = fma(0.5 * dt, rhs, basev[n]);
                    *ga1 = a1[n] + rhs;
                } else if (OP == OP_RK4_S3) {                               // :156
                    newc = fma(dt, rhs, basev[n]);
                    *ga1 = a1[n] + rhs;
                } else if (OP == OP_RK4_S4) {                               // :160
                    const double third = 1.0 / 3.0;
                    newc = fma(dt * third, fma(0.5, rhs, a1[n]), basev[n]);
                } else if (OP == OP_RK38_S1) {                              // :169
                    const double third = 1.0 / 3.0;
                    newc = fma(third * dt, rhs, c0);
                    *ga1 = rhs;
                } else if (OP == OP_RK38_S2) {                              // :174
                    const double third = 1.0 / 3.0;
                    const double r0 = a1[n];
                    newc = fma(dt, fma(-third, r0, rhs), basev[n]);
                    *ga1 = r0 - rhs;
                    *ga2 = fma(3.0, rhs, r0);
                } else if (OP == OP_RK38_S3) {                              // :178
                    newc = fma(dt, a1[n] + rhs, basev[n]);
                    *ga2 = fma(3.0, rhs, a2[n]);
                } else if (OP == OP_RK38_S4) {                              // :182
                    newc = fma(dt * 1.0 / 8.0, a2[n] + rhs, basev[n]);
                }
                if (OP != OP_RHS) {
                    raw[n] = newc;
                    if (CLIP) {
                        if (n & 1) nanf |= either_nan(raw[n - 1], raw[n]);  // component :2392
                        const double mn = p.om.minimum[n];                  // :1728-1730
                        newc = clip_min(newc, mn);
                    }
                    *go = newc;
                }
                go += plane; ga1 += plane; ga2 += plane; gr += plane;
            }
        };
        if (T::final_stage && clip_now) finish_layer(std::true_type{});
        else finish_layer(std::false_type{});
        g_out += ld; g_base += ld; g_a1 += ld; g_a2 += ld; g_rhs += ld;
    }
    cp_async_wait<0>();

    if (OP == OP_ADAPTIVE && viol) atomicOr(&p.ctl->flags[0], 1);
    if (T::final_stage && nanf) atomicOr(&p.ctl->flags[1], 1);
}
