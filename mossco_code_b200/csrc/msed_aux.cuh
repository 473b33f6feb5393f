// msed_aux.cuh -- included inside namespace msed by msed_kernels.cuh.
//
// Kernels beside the stepping path: the pelagic -> soil connector (src/mediators/pelagic_soil_connector.F90),
// whole-domain diagnostics (bed-flux sums and inventories, a position-weighted checksum of the state) and the
// fp64-pipe micro-benchmark the roofline report is measured against.

// ---- pelagic_soil_connector Run (src/mediators/pelagic_soil_connector.F90:176-2122) -------------------------
// Bottom-layer pelagic fields -> the sediment's *_at_soil_surface / *_z_velocity_at_soil_surface fields, one
// thread per column.  Rows of the output staging area [..][ld]: csurf 0..7 in the sediment's variable order
// (ldetC sdetC detP po4 no3 nh3 oxy odu), wz 0..2.
struct P2SIn {
    const double *oxygen, *odu;                 // :648-720; either may be absent
    const double *detN, *detN_wz;               // :921-1010 (required)
    const double *detC;                         // :1024-1070; absent: C:N = 106/16
    const double *detP, *detP_wz;               // :1527-1560, :1601-1645
    const double *nitrate, *ammonium, *DIN, *DIP;  // :1655-2110
    const double *depth, *tke;                  // water_depth_at_soil_surface :1126, turbulent_kinetic_energy_.. :1164
};
struct P2SPar {
    double sinking_factor, sinking_factor_min, NC_ldet, NC_sdet, half_sedimentation_depth,
        half_sedimentation_tke, critical_detritus, convertN, convertP;   // :38-46, namelist :146-148
    int head_compat;   // reproduce the HEAD revision's detritus / phosphate branches (see msed.h)
};
__global__ void pelagic_soil_kernel(double *csurf, double *wz, unsigned *written, P2SIn in, P2SPar q, size_t ld_out,
                                    int ncol)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const double detN = in.detN[col], vN = in.detN_wz[col];
    // C:N of the detritus and its split into the labile and semilabile pools, :1022-1097
    double CN = 106.0 / 16.0;
    if (in.detC) CN = __ddiv_rn(in.detC[col], __dadd_rn((double)1E-5f, detN));      // :1063 (1E-5 is default real)
    double fac_l = __ddiv_rn(__dsub_rn(1.0, __dmul_rn(q.NC_sdet, CN)), __dsub_rn(q.NC_ldet, q.NC_sdet));  // :1082
    if (fac_l > CN) fac_l = CN;                                                      // :1084-1087
    if (fac_l < 0.0) fac_l = 0.0;                                                    // :1088-1091
    const double fac_s = __dsub_rn(CN, fac_l);                                       // :1095
    // environmental sinking factor, :1097-1232
    double fac_env = 1.0;
    if (in.depth && q.half_sedimentation_depth > (double)1E-3f) {                    // :1150-1155
        const double d2 = __dmul_rn(in.depth[col], in.depth[col]);
        fac_env = __ddiv_rn(__dmul_rn(fac_env, d2),
                            __dadd_rn(d2, __dmul_rn(q.half_sedimentation_depth, q.half_sedimentation_depth)));
    }
    if (in.tke && q.half_sedimentation_tke < (double)9E9f)                           // :1189-1191
        fac_env = __ddiv_rn(__dmul_rn(fac_env, q.half_sedimentation_tke),
                            __dadd_rn(in.tke[col], q.half_sedimentation_tke));
    fac_env = __dadd_rn(fac_env, __ddiv_rn(q.sinking_factor_min, q.sinking_factor)); // :1204
    if (in.detC && q.critical_detritus > (double)1E-3f && q.critical_detritus < (double)9E9f) {  // :1215-1227
        const double x = __ddiv_rn(in.detC[col], q.critical_detritus);
        const double x2 = __dmul_rn(x, x);                                           // x**4 = (x*x)*(x*x)
        fac_env = __ddiv_rn(__dmul_rn(fac_env, 1.0), __dadd_rn(1.0, __dmul_rn(x2, x2)));
    }
    const double sink = __dmul_rn(q.sinking_factor, fac_env);
    if (!q.head_compat) {
        csurf[0 * ld_out + col] = __dmul_rn(__dmul_rn(fac_l, q.convertN), detN);     // :1282-1284
        csurf[1 * ld_out + col] = __dmul_rn(__dmul_rn(fac_s, q.convertN), detN);     // :1342-1344
        wz[0 * ld_out + col] = __dmul_rn(sink, vN);     // the velocity fields get sinking_factor*fac_env*velocity
        wz[1 * ld_out + col] = __dmul_rn(sink, vN);
        wz[2 * ld_out + col] = __dmul_rn(sink, in.detP_wz ? in.detP_wz[col] : vN);
    } else {
        // HEAD fetches fieldList(1) -- the concentration field -- a second time "for velocity field" and
        // overwrites it with sinking_factor*fac_env*detN (:1291-1295, :1351-1355); the two carbon velocity
        // fields are never written.  The phosphorus velocity gets the same expression, with detN where a
        // velocity was meant, unless a detP velocity is imported (:1595-1597, :1626-1630)
        csurf[0 * ld_out + col] = __dmul_rn(sink, detN);
        csurf[1 * ld_out + col] = __dmul_rn(sink, detN);
        wz[2 * ld_out + col] = __dmul_rn(sink, in.detP_wz ? in.detP_wz[col] : detN);
    }
    csurf[2 * ld_out + col] = in.detP ? in.detP[col]                                  // :1557-1559
                                      : __dmul_rn(__dmul_rn(1.0 / 16.0, q.convertN), detN);  // :1521-1523
    // nutrients, :1775-1965
    const bool hasN = in.nitrate != nullptr, hasA = in.ammonium != nullptr, hasD = in.DIN != nullptr;
    const double nit = hasN ? in.nitrate[col] : 0.0, amm = hasA ? in.ammonium[col] : 0.0, din = hasD ? in.DIN[col] : 0.0;
    double a_out, n_out;
    if (hasA) a_out = __dmul_rn(q.convertN, amm);                                     // :1816-1818
    else if (hasD && hasN) a_out = __dmul_rn(q.convertN, __dsub_rn(din, nit));        // :1821-1823
    else if (hasD) a_out = __dmul_rn(__dmul_rn(q.convertN, 0.5), din);                // :1832-1834
    else a_out = __dmul_rn(q.convertN, nit);                                          // :1843-1845
    if (hasN) n_out = __dmul_rn(q.convertN, nit);                                     // :1930-1932
    else if (hasA && hasD) n_out = __dmul_rn(q.convertN, __dsub_rn(din, amm));        // :1939-1943
    else if (hasD) n_out = __dmul_rn(__dmul_rn(q.convertN, 0.5), din);                // :1952-1954
    else n_out = __dmul_rn(q.convertN, amm);                                          // :1963-1965
    csurf[4 * ld_out + col] = n_out;
    csurf[5 * ld_out + col] = a_out;
    // phosphate, :1985-2110: DIP if imported, else Redfield from DIN (built from nitrate + ammonium, or twice
    // the one that exists, when DIN itself is absent :2040-2070).  HEAD recomputes dip from DIN even when DIP
    // was found (:2092-2094), overwriting the imported field.
    double din_eff = din;
    if (!hasD) din_eff = (hasA && hasN) ? __dadd_rn(nit, amm) : (hasA ? __dmul_rn(2.0, amm) : __dmul_rn(2.0, nit));
    const bool din_known = hasD || hasA || hasN;
    if (in.DIP && !(q.head_compat && din_known))
        csurf[3 * ld_out + col] = __dmul_rn(q.convertP, in.DIP[col]);
    else
        csurf[3 * ld_out + col] = __dmul_rn(q.convertP, __dmul_rn(__dmul_rn(1.0 / 16.0, q.convertN), din_eff));
    // oxygen and reduced substances, :800-860
    if (in.oxygen && in.odu) {                       // both imported: plain copies (:806-807)
        csurf[6 * ld_out + col] = in.oxygen[col];
        csurf[7 * ld_out + col] = in.odu[col];
    } else if (in.odu) {                             // only odu: its negative part is oxygen (:827-828)
        const double o = in.odu[col];
        csurf[6 * ld_out + col] = -o > 0.0 ? -o : 0.0;
        csurf[7 * ld_out + col] = o > 0.0 ? o : 0.0;
    } else if (in.oxygen) {                          // only oxygen: its negative part is odu (the intent of
        const double o = in.oxygen[col];             // :849-850, which at HEAD dereferences the odu pointer)
        csurf[6 * ld_out + col] = o > 0.0 ? o : 0.0;
        csurf[7 * ld_out + col] = -o > 0.0 ? -o : 0.0;
    }
    (void)written;
}

// ---- whole-domain diagnostics --------------------------------------------------------------------------------
// Position-weighted checksum of the state over the wet columns: sum over cells of bits(conc) * (2*g + 1) mod 2^64
// and the XOR of bits(conc), g = the cell's index in the GLOBAL Fortran array (j-slab tiles: col_offset =
// j_offset*inum).  Both are independent of how the domain is cut into tiles, so a sharded run must reproduce the
// single-tile values (bench.py prints them).
__global__ void checksum_kernel(const double *state, const unsigned char *mask, size_t ld, int ncol, int rows,
                                long long global_ncol, long long col_offset, unsigned long long *out)
{
    unsigned long long sum = 0, x = 0;
    const long long total = (long long)rows * ncol;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / ncol, col = t - row * ncol;
        if (mask[col]) continue;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(state[(size_t)row * ld + col]);
        const unsigned long long g = (unsigned long long)(row * global_ncol + col_offset + col);
        sum += bits * (2ull * g + 1ull);
        x ^= bits;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], sum);
        atomicXor(&out[1], x);
    }
}

// per-variable sums over the wet columns of the tile: bed flux (fluxes(:,:,n), mmol m-2 s-1 summed over columns)
// and inventory (sum_k conc*porosity*dz, mmol m-2 summed over columns).  Two passes with a fixed summation order:
// block partials [gridDim.x][2*NV], then one block adds them up in index order.
__global__ void diag_partial_kernel(double *partial, const double *state, const double *por, const double *fluxes,
                                    const unsigned char *mask, size_t ld, int ncol, int K, KParams p)
{
    __shared__ double sh[256];
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const bool wet = col < ncol && !mask[col];
    for (int n = 0; n < 2 * NV; ++n) {
        double v = 0.0;
        if (wet) {
            if (n < NV) {
                v = fluxes[(size_t)n * ld + col];
            } else {
                const int m = n - NV;
                for (int k = 0; k < K; ++k)
                    v += state[(size_t)(m * K + k) * ld + col] * por[(size_t)k * ld + col] * p.dz[k];
            }
        }
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) partial[(size_t)blockIdx.x * 2 * NV + n] = sh[0];
        __syncthreads();
    }
}
__global__ void diag_final_kernel(double *out, const double *partial, int nblocks)
{
    const int n = threadIdx.x;
    if (n >= 2 * NV) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * 2 * NV + n];
    out[n] = s;
}

// cross-tile minloc of the adaptive solver diagnostics (solver_library.F90:131-135): each tile contributes the
// relative change and the GLOBAL Fortran index of its own minloc; the smallest value wins, the smallest index
// among equals (Fortran's minloc returns the first).  pack: [0] value as double, [1] index as int64.
__global__ void minloc_pack_kernel(double *val_out, long long *idx_out, const double *best_val,
                                   const long long *best_idx, int ncol, long long global_ncol, long long col_offset)
{
    const long long idx = *best_idx;
    if (idx < 0) {
        *val_out = 1.0e300;
        *idx_out = 0x7fffffffffffffffLL;
        return;
    }
    const long long row = idx / ncol, col = idx % ncol;
    *val_out = *best_val;
    *idx_out = row * global_ncol + col_offset + col;
}
__global__ void minloc_select_kernel(long long *idx, const double *my_val, const double *min_val)
{
    if (*my_val != *min_val) *idx = 0x7fffffffffffffffLL;   // only tiles that hold the minimum compete for the index
}

// ---- fp64 pipe micro-benchmark (what roofline.peak of a fused, fp64-bound launch is measured with) -----------
__global__ void dfma_peak_kernel(double *out, double a, double b, int iters)
{
    double x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;
}
