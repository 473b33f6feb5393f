// msed.cu -- host side and C ABI (include/msed.h) of the B200-native fabm_sediment column solver.
//
// Host-side restatement of type_sed (src/drivers/fabm_sediment_driver.F90:69-113) and of the
// scalar control flow of ode_solver (src/utilities/solver_library.F90:80-189): the arrays live on
// the device in the reference's own Fortran layout, the column kernels in msed_kernels.cuh do the
// array work.  There is no CPU compute path in this file.
#include "msed_kernels.cuh"
#include "msed_launch.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <new>
#include <string>
#include <vector>

using namespace msed;

// ---- NCCL bound at run time (no link-time dependency) -----------------------------------------
namespace {
typedef struct ncclComm *ncclComm_t;
struct NcclUniqueId { char internal[128]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi &nccl_api()
{
    static NcclApi api;
    if (api.lib) return api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return api;
    api.GetUniqueId = (int (*)(NcclUniqueId *))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t *, int, NcclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(
        api.lib, "ncclAllReduce");
    api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce;
    return api;
}
constexpr int kNcclInt32 = 2;  // ncclInt32
constexpr int kNcclInt64 = 4;  // ncclInt64
constexpr int kNcclFloat64 = 8;  // ncclFloat64
constexpr int kNcclSum = 0;    // ncclSum
constexpr int kNcclMax = 2;    // ncclMax
constexpr int kNcclMin = 3;    // ncclMin
}  // namespace

// ---- handle -----------------------------------------------------------------------------------
struct msed_handle {
    msed_config cfg;
    int K = 0, ncol = 0;
    size_t ld = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    double *buf[2] = {nullptr, nullptr};
    double *aux[2] = {nullptr, nullptr};
    double *por = nullptr, *bdys = nullptr, *fluxes = nullptr, *par_surface = nullptr;
    double *scratch = nullptr;  // [nvar][K][ld] staging (rhs / fields / packed transfers)
    double *denit = nullptr;    // [K][ld] denit diagnostic left by the last fused pair of a call
    bool denit_valid = false;   // ... and whether it describes the current state
    double *pel = nullptr;      // pelagic boxes: conc [nvar][ld], wz [nvar][ld], height [ld], temperature [ld]
    double *tables = nullptr;   // device copies of zc[K], cumdepth[K], porosity profile[K]
    unsigned char *mask = nullptr;
    int *colmap = nullptr;          // device list of the wet columns (null while the tile has no land)
    std::vector<int> wet_idx;       // host copy: converts column ranges of chunked launches
    bool has_land = false;          // msed_set_mask marked at least one column as land
    Ctl *ctl = nullptr;         // device
    Ctl *ctl_host = nullptr;    // pinned mirror
    double *minloc_val = nullptr;
    long long *minloc_idx = nullptr;
    double *red = nullptr;      // 32 doubles of device scratch for the small cross-tile reductions
    double *snap = nullptr;     // [nvar][K][ld] snapshot of the state an asynchronous export reads (msed_export_state_begin)
    cudaEvent_t ev_snap = nullptr, ev_export = nullptr;
    cudaStream_t export_stream = nullptr;   // its own stream: the per-Run flux copies on copy_stream must not queue behind it
    bool export_pending = false, export_direct = false;
    // the export is handed to the copy engine a few column slices at a time (export_pump): a device-to-host engine
    // works its queue off in submission order, and a state's worth of slices submitted at once would hold up the
    // bed-flux copies of every Run until the whole export is through
    const double *export_src = nullptr;
    double *export_dst = nullptr;
    size_t export_next = 0, export_slice = 0;    // next element of the dense host array to submit; elements per slice
    bool export_submitted = false;               // every slice is in the queue and ev_export is recorded
    int compat = 0;             // MSED_COMPAT_* (msed_set_compat)
    int cur = 0;
    int por_mode = 1;           // how the column kernel obtains porosity (see KParams::por_mode)
    std::vector<double> zi, zc, dz, dzc, bf, por_profile, cumdepth;
    double bioturbation_eff = 0.0;  // sed%bioturbation (profile 3 overwrites it with 1, driver :623)
    double last_min_dt = (double)1.e20f;  // solver_library.F90:44 (default-real literal)
    int last_min_dt_grid_cell[4] = {-99, -99, -99, -99};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;
    cudaStream_t copy_stream = nullptr;           // msed_run_exchange: H2D of the import fields
    cudaStream_t stream2 = nullptr;               // chunk-major Run: odd chunks run here, so that the tail of one chunk's
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // launches overlaps the head of the next chunk's
    cudaStream_t d2h_stream = nullptr;            // ... D2H of the bed fluxes, on its own stream: the fluxes of chunk c
                                                  // leave while the import fields of later chunks still arrive
    cudaEvent_t ev_pool[2 * 16] = {};             // per-chunk H2D-done / compute-done events
    cudaEvent_t ev_x[3] = {};                     // msed_run_exchange phase marks: first H2D issued, last H2D landed, last D2H landed
    double exchange_ms[4] = {0, 0, 0, 0};         // msed_get_exchange_timing of the last pipelined msed_run_exchange
    int exchange_chunks = 0;                      // 0 = choose from the tile size
    int step_fusion = 1;                          // 0 off, 1 auto (chains where they apply, else pairs),
                                                  // 2 pairs only, 3 chains wherever knum allows
    long long chain_max_cols = 0;                 // auto mode: tiles up to this many columns take chain_kernel
    int rk_stages = 4;                            // Runge-Kutta, thread per column: stages per launch (4: rk_quad_kernel, 2: rk_pair_kernel)
    double *xstage = nullptr;                     // [20][ld] staging rows of msed_run_exchange when the staging buffer
                                                  // itself serves as third state buffer (chunk-major Run)
    int chunk_major = 1;                          // msed_run_exchange: whole coupling interval chunk by chunk
    int pred_depth = 0;                           // how the next call's steps are predicted to go: every step at
                                                  // dt/4^pred_depth (what the last completed step did); -1: no prediction
    int regime_depth = 0;                         // depth of the last fused group that was committed: the prediction
                                                  // after a step whose own attempt sequence was irregular
    long long pairs_committed = 0;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    msed_allreduce_hook hook = nullptr;
    void *hook_user = nullptr;
    // msed_set_import_generations: what the staging rows of msed_run_exchange hold from earlier Runs
    bool import_gen_on = false;
    unsigned long long import_gen[1 + 2 * NV] = {};        // the caller's current counters
    struct ImportRow { const double *host = nullptr; unsigned long long gen = 0; int row = -1; const double *stage = nullptr; };
    ImportRow import_row[1 + 2 * NV];                       // last upload of every field
    // TMA descriptors of the state-sized buffers pair_kernel reads (launch_pair), keyed by base pointer
    std::vector<std::pair<const void *, CUtensorMap>> tmaps;
    std::string err;
};

namespace {

thread_local std::string g_err;

int fail(msed_handle *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    g_err = msg;
    return code;
}

#define CUDA_TRY(h, expr)                                                                        \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(h, MSED_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(e_));     \
    } while (0)

inline int nblocks(int ncol, int bs = 256) { return (ncol + bs - 1) / bs; }

// host <-> device copies between the caller's dense Fortran arrays (row length ncol) and the
// padded device planes (row length ld)
int upload_rows(msed_handle *h, double *dst, const double *src, size_t rows)
{
    CUDA_TRY(h, cudaMemcpy2DAsync(dst, h->ld * sizeof(double), src, (size_t)h->ncol * sizeof(double),
                                  (size_t)h->ncol * sizeof(double), rows, cudaMemcpyHostToDevice,
                                  h->stream));
    return MSED_OK;
}
int download_rows(msed_handle *h, double *dst, const double *src, size_t rows)
{
    CUDA_TRY(h, cudaMemcpy2DAsync(dst, (size_t)h->ncol * sizeof(double), src, h->ld * sizeof(double),
                                  (size_t)h->ncol * sizeof(double), rows, cudaMemcpyDeviceToHost,
                                  h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

// hzg_omexdia_p namelist values (rates per day) -> what the kernels use (rates per second, folded constants)
OmexDev omex_dev(const msed_config &c)
{
    OmexDev m;
    m.rLabile = c.rLabile / 86400.0;
    m.rSemilabile = c.rSemilabile / 86400.0;
    m.NCrLdet = c.NCrLdet;
    m.NCrSdet = c.NCrSdet;
    m.PAds_rS = c.PAds * m.rSemilabile;
    m.PAdsODU = c.PAdsODU;
    m.rNH3Ads = 1.0 / (1.0 + c.NH3Ads);
    m.CprodMax = c.CprodMax / 86400.0;
    m.rnit = c.rnit / 86400.0;
    m.ksO2nitri = c.ksO2nitri;
    m.rODUox = c.rODUox / 86400.0;
    m.ksO2oduox = c.ksO2oduox;
    m.ksO2oxic = c.ksO2oxic;
    m.ksNO3denit = c.ksNO3denit;
    m.kinO2denit = c.kinO2denit;
    m.kinNO3anox = c.kinNO3anox;
    m.kinO2anox = c.kinO2anox;
    m.E_a = 0.1 * std::log(1.5) * 288.15 * (288.15 + 10.0);
    for (int n = 0; n < NV; ++n) m.minimum[n] = c.minimum[n];
    return m;
}

void fill_params(const msed_handle *h, KParams &p)
{
    std::memset(&p, 0, sizeof(p));
    const msed_config &c = h->cfg;
    p.buf[0] = h->buf[0];
    p.buf[1] = h->buf[1];
    p.aux1 = h->aux[0];
    p.aux2 = h->aux[1];
    p.rhs_out = h->scratch;
    p.denit_out = nullptr;
    p.por = h->por;
    p.bdys = h->bdys;
    p.fluxes = h->fluxes;
    p.mask = h->mask;
    p.colmap = nullptr;
    p.feed_bulk = 0;
    p.min_zero = 1;
    for (int n = 0; n < NV; ++n)
        if (h->cfg.minimum[n] != 0.0) p.min_zero = 0;
    p.ctl = h->ctl;
    p.ld = h->ld;
    p.ncol = h->ncol;
    p.col0 = 0;
    p.col_end = h->ncol;
    p.K = h->K;
    p.inum = c.inum;
    p.i_offset = c.i_offset;
    p.j_offset = c.j_offset;
    p.bcup_diss = c.bcup_dissolved_variables;
    p.bcup_part = c.distributed_pom_flux ? 4 : 1;  // driver :239-243
    p.profile = c.bioturbation_profile;
    p.use_ctl = 1;
    p.por_mode = h->por_mode;
    p.dt = 0.0;
    p.fac = 1.0 + c.relative_change_min;           // solver_library.F90:121
    p.bioturbation = h->bioturbation_eff;
    p.diffusivity = c.diffusivity;
    p.pom_flux_rate = c.pom_flux_max / 86400.0;     // driver :284
    p.beta = c.bioturb_beta;
    p.b = c.bioturb_b;
    p.L1 = c.bioturb_L1;
    p.L2 = c.bioturb_L2;
    p.poc_factor[0] = 1.0 / 1.2 * 12.01 / c.bioturb_dry_density / 1000.0;  // driver :383
    p.poc_factor[1] = 1.0 / 6.0 * 12.01 / c.bioturb_dry_density / 1000.0;  // driver :386
    p.cumdepth_last = h->cumdepth[h->K - 1];
    p.om = omex_dev(c);
    for (int k = 0; k < h->K; ++k) {
        p.dz[k] = h->dz[k];
        p.rdzc[k] = (k < h->K - 1) ? 1.0 / h->dzc[k] : 0.0;
        p.bf[k] = h->bf[k];
        p.e1[k] = std::exp(h->zc[k] * 100.0 * c.bioturb_k_l);  // driver :639
        p.e2[k] = std::exp(h->zc[k] * 200.0 * c.bioturb_k_l);  // driver :640
        p.portab[k] = (h->por_mode == 2) ? 1.0 - c.porosity_fac * (h->zc[k] - h->zc[0])  // driver :411-412
                                         : h->por_profile[k];                                // driver :280
    }
}

cudaError_t launch_column(const msed_handle *h, int op, const KParams &p)
{
    // the (rare) Zhang & Wirtz path always streams the 3-D porosity field, which is kept current in
    // every mode; the closed-form porosity variants exist for the hot profile-0/1/2 kernels only
    const bool p3 = h->cfg.bioturbation_profile == 3;
    const bool sp = p3 || p.por_mode == 0;
    return tu_launch_column(h->cfg.model, p3, sp, op, p, h->stream);
}

int ensure_aux(msed_handle *h, int count)
{
    const size_t bytes = (size_t)NV * h->K * h->ld * sizeof(double);
    for (int a = 0; a < count; ++a)
        if (!h->aux[a]) {
            CUDA_TRY(h, cudaMalloc(&h->aux[a], bytes));
            CUDA_TRY(h, cudaMemsetAsync(h->aux[a], 0, bytes, h->stream));
        }
    return MSED_OK;
}

int ensure_scratch(msed_handle *h)
{
    if (!h->scratch) {
        const size_t bytes = (size_t)NV * h->K * h->ld * sizeof(double);
        CUDA_TRY(h, cudaMalloc(&h->scratch, bytes));
    }
    return MSED_OK;
}

int ensure_denit(msed_handle *h)
{
    if (!h->denit) CUDA_TRY(h, cudaMalloc(&h->denit, (size_t)h->K * h->ld * sizeof(double)));
    return MSED_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// descriptor of one state buffer [nvar][K][ld] for pair_kernel's box copies; false if it cannot be made
bool state_tmap(msed_handle *h, const double *base, CUtensorMap *out)
{
    for (const auto &e : h->tmaps)
        if (e.first == base) { *out = e.second; return true; }
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)h->ld, (cuuint64_t)h->K, (cuuint64_t)NV};
    const cuuint64_t strides[2] = {(cuuint64_t)h->ld * 8, (cuuint64_t)h->ld * 8 * (cuuint64_t)h->K};
    const cuuint32_t box[3] = {32, 1, (cuuint32_t)NV};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (h->tmaps.size() >= 8) h->tmaps.erase(h->tmaps.begin());
    h->tmaps.emplace_back(base, m);
    *out = m;
    return true;
}

cudaError_t launch_pair(msed_handle *h, int method, const KParams &pin, cudaStream_t stream = nullptr)
{
    KParams p = pin;
    if (h->colmap) {  // masked tile: run over the wet columns of [col0, col_end) only
        const auto lo = std::lower_bound(h->wet_idx.begin(), h->wet_idx.end(), pin.col0);
        const auto hi = std::lower_bound(h->wet_idx.begin(), h->wet_idx.end(), pin.col_end);
        p.col0 = (int)(lo - h->wet_idx.begin());
        p.col_end = (int)(hi - h->wet_idx.begin());
        p.colmap = h->colmap;
        if (p.col_end <= p.col0) return cudaSuccess;  // all land: nothing to launch
    }
    p.feed_bulk = 0;
    if (!h->has_land && !h->colmap) {
        // opt-in: measured slower than the per-thread cp.async ring on B200 (profiles/r02_summary.md)
        static const bool off = !(std::getenv("MSED_PAIR_FEED") && !std::strcmp(std::getenv("MSED_PAIR_FEED"), "bulk"));
        if (!off && state_tmap(h, p.buf[0], &p.tmap[0]) && state_tmap(h, p.buf[1], &p.tmap[1]) &&
            (!p.in_ovr || state_tmap(h, p.in_ovr, &p.tmap[2])))
            p.feed_bulk = 1;
    }
    return tu_launch_pair(h->cfg.model, method == MSED_ADAPTIVE_EULER, p, stream ? stream : h->stream);
}

// one launch = m ode_solver calls, warp per column (msed_chain.cuh)
cudaError_t launch_chain(const msed_handle *h, int method, const KParams &p, int m, bool clip)
{
    return tu_launch_chain(h->cfg.model, method == MSED_ADAPTIVE_EULER, clip, p, m, h->stream);
}

// one launch = m Runge-Kutta ode_solver calls, warp per column (msed_chain.cuh)
cudaError_t launch_rk_chain(const msed_handle *h, int method, const KParams &p, int m, bool clip)
{
    return tu_launch_rk_chain(h->cfg.model, method, clip, p, m, h->stream);
}

// one launch = the four stages of a Runge-Kutta call (msed_rkquad.cuh)
cudaError_t launch_rk_quad(const msed_handle *h, int method, const KParams &pin)
{
    KParams p = pin;
    if (h->colmap) {  // masked tile: run over the wet columns of [col0, col_end) only (as launch_pair)
        const auto lo = std::lower_bound(h->wet_idx.begin(), h->wet_idx.end(), pin.col0);
        const auto hi = std::lower_bound(h->wet_idx.begin(), h->wet_idx.end(), pin.col_end);
        p.col0 = (int)(lo - h->wet_idx.begin());
        p.col_end = (int)(hi - h->wet_idx.begin());
        p.colmap = h->colmap;
        if (p.col_end <= p.col0) return cudaSuccess;  // all land: nothing to launch
    }
    return tu_launch_rk_quad(h->cfg.model, method, p, h->stream);
}

// one launch = two chained Runge-Kutta stages (msed_rkpair.cuh); which = 0 for stages 1+2, 1 for 3+4
cudaError_t launch_rk_pair(const msed_handle *h, int method, int which, const KParams &p)
{
    return tu_launch_rk_pair(h->cfg.model, method, which, p, h->stream);
}

cudaError_t enable_pair_smem()
{
    cudaError_t e = tu_enable_pair_smem();
    if (e == cudaSuccess) e = tu_enable_rk_smem();
    return e != cudaSuccess ? e : tu_enable_rk_quad_smem();
}

// nflags: 4 (violation / NaN of up to two stages) or, for a group that plans rejections, all of Ctl::flags
int reduce_flags(msed_handle *h, int nflags = 8)
{
    if (h->hook) {
        void *flags = (void *)((char *)h->ctl + offsetof(Ctl, flags));
        if (h->hook(h->hook_user, flags, nflags, (void *)h->stream) != 0)
            return fail(h, MSED_ERR_NCCL, "allreduce hook failed");
    } else if (h->comm) {
        NcclApi &api = nccl_api();
        int *flags = (int *)((char *)h->ctl + offsetof(Ctl, flags));
        int rc = api.AllReduce(flags, flags, (size_t)nflags, kNcclInt32, kNcclMax, h->comm, h->stream);
        if (rc != 0)
            return fail(h, MSED_ERR_NCCL, std::string("ncclAllReduce: ") +
                                              (api.GetErrorString ? api.GetErrorString(rc) : "error"));
    }
    return MSED_OK;
}

// hand the next slices of a pending state export to the copy engine: about `budget_bytes` worth (everything if < 0).
// Slices are plain contiguous copies -- runs of whole rows when the device rows are unpadded, pieces of one row
// otherwise -- of at most export_slice elements; export_next counts elements of the dense host array.
int export_pump(msed_handle *h, double budget_bytes)
{
    if (!h->export_pending || h->export_submitted) return MSED_OK;
    const size_t ncol = (size_t)h->ncol, total = (size_t)NV * h->K * ncol;
    const bool dense = h->ld == ncol;
    double sent = 0.0;
    while (h->export_next < total && (budget_bytes < 0 || sent < budget_bytes)) {
        const size_t e = h->export_next, r = e / ncol, c = e % ncol;
        size_t n = std::min(h->export_slice, total - e);
        if (!dense) n = std::min(n, ncol - c);           // stay inside the row
        CUDA_TRY(h, cudaMemcpyAsync(h->export_dst + e, h->export_src + r * h->ld + c, n * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->export_stream));
        h->export_next += n;
        sent += (double)(n * sizeof(double));
    }
    if (h->export_next >= total) {
        CUDA_TRY(h, cudaEventRecord(h->ev_export, h->export_stream));
        h->export_submitted = true;
    }
    return MSED_OK;
}

// bytes of a pending export to hand over after a stepping call whose kernels took kernel_ms: what the engine moves in
// half of that time, but at least a sixteenth of the state -- the export is through after sixteen calls at the latest
// (an output cadence is tens of Runs), also where a Run is short against its tile's state (C3: 3 ms, 1.9 GB)
double export_helping(const msed_handle *h, double kernel_ms)
{
    const double state_bytes = (double)NV * h->K * (double)h->ncol * sizeof(double);
    return std::max(std::max(16.0e6, state_bytes / 16.0), 0.5 * kernel_ms * 1.0e-3 * 50.0e9);
}

// msed_run_exchange: the tile is cut into column chunks so that the H2D of the import fields overlaps
// the first attempt and the D2H of the bed fluxes overlaps the last one
struct ExchangePlan {
    int nchunks = 0;
    int c0[16], c1[16];
    BcPtrs bc;                 // device staging rows of the import fields
    bool first = false;        // chunk the first attempt: wait for chunk H2D, assemble boundary, compute
    bool last = false;         // chunk the last attempt: export each chunk as soon as it is computed
    double *neg = nullptr;     // device rows [nvar][ld] receiving -fluxes
    double *host_out = nullptr;
    bool export_done = false;  // out: host_out holds the final upward fluxes
    bool stage_private = false;  // the staging rows are not in h->scratch, which may then hold a state
    bool zero_copy = false;      // small tile, pinned host fields: the boundary kernel reads the import fields and the
    double *host_out_dev = nullptr;  // export kernel writes the fluxes straight through PCIe (no copies, no staging)
};

// One fused launch of a call's plan: a pair (two accepted sub-steps), or a chain of m ode_solver calls
struct FusedLaunch {
    int kind = PAIR_FULL;   // PairKind; chains ignore it
    int m = 0;              // chain: ode_solver calls in the launch
    int step = 0;           // index (within the call) of the ode_solver call the launch starts in
    PlanCommit pc;          // what committing this launch alone does to the control block
};

// The plan of a call: every one of its nsteps ode_solver calls is predicted to run at dt/4^depth -- the way
// the last completed call went (Ctl::last_depth; sub-cycling comes in episodes of many consecutive steps) --
// i.e. as the reference's attempt sequence  dt (rejected) .. dt/4^(depth-1) (rejected), then 4^depth accepted
// sub-steps of dt_acc = dt/4^depth (solver_library.F90:104-140).  Returns false if that sequence cannot be
// planned: a step size that must be rejected is not rejectable (:126), or the sub-steps do not add up to dt in
// the reference's own floating-point loop (:108,:138).
bool plan_depth(double dt, double dt_min, int depth, double &dt_acc, long long &nq)
{
    if (depth < 0 || depth > MAX_PLAN_DEPTH) return false;
    double dq = dt;
    for (int l = 0; l < depth; ++l) {
        if (!(dq > dt_min)) return false;   // a violation at dq would be accepted as it is
        dq *= 0.25;                          // :127
    }
    nq = 1LL << (2 * depth);
    double di = 0.0;
    for (long long i = 1; i <= nq; ++i) {
        di = di + dq;                        // :138
        if ((i < nq) != (di < dt)) return false;
    }
    dt_acc = dq;
    return true;
}

// the step loop shared by msed_ode_solver / msed_step / msed_run / msed_run_exchange
int run_steps(msed_handle *h, double dt, int method, long long nsteps, bool wrapper, msed_step_info *info,
              ExchangePlan *plan = nullptr)
{
    if (nsteps < 0) return fail(h, MSED_ERR_ARG, "nsteps < 0");
    if (method < 0 || method > 3) return fail(h, MSED_ERR_ARG, "unknown ode_method");
    if (h->cfg.model == MSED_MODEL_TEST_SOLVER && method != MSED_EULER)
        return fail(h, MSED_ERR_ARG, "MSED_MODEL_TEST_SOLVER supports MSED_EULER only");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const bool diag = h->cfg.adaptive_solver_diagnostics && method == MSED_ADAPTIVE_EULER;
    int rc;
    if (method == MSED_RUNGE_KUTTA_4) { if ((rc = ensure_aux(h, 1))) return rc; }
    if (method == MSED_RUNGE_KUTTA_4_38) { if ((rc = ensure_aux(h, 2))) return rc; }

    if (h->cfg.bioturbation_profile == 3) h->bioturbation_eff = 1.0;  // driver :623

    Ctl c;
    std::memset(&c, 0, sizeof(c));
    c.dt = dt;
    c.dt_int = 0.0;
    c.dt_red = dt;
    c.dt_min = h->cfg.dt_min;
    c.last_min_dt = h->last_min_dt;
    c.steps_done = 0;
    c.steps_target = nsteps;
    c.cur = h->cur;
    c.do_clip = (wrapper && !diag) ? 1 : 0;
    c.diagnostics = diag ? 1 : 0;
    c.last_depth = h->pred_depth < 0 ? 0 : h->pred_depth;
    *h->ctl_host = c;
    CUDA_TRY(h, cudaMemcpyAsync(h->ctl, h->ctl_host, sizeof(Ctl), cudaMemcpyHostToDevice, h->stream));

    const bool wrapper_clip = c.do_clip != 0;  // check_NaN + minimum clip after every step (chain_kernel<.., CLIP>)
    KParams p;
    fill_params(h, p);
    InitVals minimum;
    for (int n = 0; n < NV; ++n) minimum.v[n] = h->cfg.minimum[n];
    // tiles that share an accept decision must issue the same sequence of flag reductions
    const bool collective = (h->comm != nullptr || h->hook != nullptr);
    long long launches = 0;

    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));

    // get_boundary_conditions of one chunk of msed_run_exchange as soon as its fields have landed
    auto boundary_chunk = [&](int c, cudaStream_t st = nullptr) -> int {
        if (!st) st = h->stream;
        const int c0 = plan->c0[c], c1 = plan->c1[c];
        CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_pool[c], 0));
        BcPtrs bc = plan->bc;
        if (bc.temperature) bc.temperature += c0;
        for (int n = 0; n < NV; ++n) {
            if (bc.csurf[n]) bc.csurf[n] += c0;
            if (bc.wz[n]) bc.wz[n] += c0;
        }
        boundary_kernel<<<nblocks(c1 - c0), 256, 0, st>>>(
            h->bdys + c0, h->fluxes + c0, h->buf[h->cur] + c0, h->por + c0, bc, h->ld, h->ld, c1 - c0, h->K,
            h->cfg.bcup_dissolved_variables, h->bioturbation_eff, h->cfg.diffusivity, h->dz[0]);
        launches += 1;
        return MSED_OK;
    };

    // ---- the fused part of the call (msed_pair.cuh, msed_chain.cuh): speculative, nothing is committed on
    // ---- failure ---------------------------------------------------------------------------------------
    const bool single_attempt = (method == MSED_EULER || method == MSED_ADAPTIVE_EULER);
    // the configurations the fused kernels cover (msed_pair.cuh, msed_rkpair.cuh); the rest takes
    // one launch per attempt / per stage
    const bool fusable = h->step_fusion &&
                         (h->cfg.model == MSED_MODEL_OMEXDIA_P || h->cfg.model == MSED_MODEL_NONE) &&
                         h->cfg.bioturbation_profile != 3 && !h->cfg.distributed_pom_flux && h->por_mode != 0;
    const bool rk_fused = fusable && !single_attempt;
    h->denit_valid = false;
    // both fused kernels only work on wet columns, so the tile size that decides between them counts those
    const long long work_cols = h->colmap ? (long long)h->wet_idx.size() : (long long)h->ncol;
    // (two layers per lane above 32 layers: one CTA per SM, so the thread-per-column pairs catch up earlier)
    // (measured at K = 40: 15 vs 19 us per step at 10,000 columns, even at 16,384, 42 vs 29 at 32,761)
    const long long chain_cols = h->K <= 32 ? h->chain_max_cols : h->chain_max_cols / 5;
    const bool chain_fit = h->K <= TU_CHAIN_MAX_LAYERS &&
                           (h->step_fusion == 3 ||
                            (h->step_fusion == 1 && !collective && work_cols <= chain_cols));
    // (with a collective every rank has to take the same decision, and the tile size is a per-rank fact:
    //  auto mode then stays with pairs; mode 3, set on every rank, selects chains)

    // export of one chunk of msed_run_exchange while the next chunks are still being computed
    auto export_chunk = [&](int c, cudaStream_t st = nullptr) -> int {
        if (!st) st = h->stream;
        const int c0 = plan->c0[c], c1 = plan->c1[c];
        CUDA_TRY(h, cudaEventRecord(h->ev_pool[16 + c], st));
        CUDA_TRY(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_pool[16 + c], 0));
        if (plan->zero_copy) {   // -fluxes straight into the caller's pinned array (rows ncol apart)
            negate_rows_kernel<<<nblocks(c1 - c0), 256, 0, h->d2h_stream>>>(plan->host_out_dev + c0, h->fluxes + c0,
                                                                             (size_t)h->ncol, h->ld, c1 - c0, NV);
            launches += 1;
            return MSED_OK;
        }
        negate_rows_kernel<<<nblocks(c1 - c0), 256, 0, h->d2h_stream>>>(plan->neg + c0, h->fluxes + c0, h->ld, h->ld,
                                                                         c1 - c0, NV);
        launches += 1;
        CUDA_TRY(h, cudaMemcpy2DAsync(plan->host_out + c0, (size_t)h->ncol * sizeof(double), plan->neg + c0,
                                      h->ld * sizeof(double), (size_t)(c1 - c0) * sizeof(double), NV,
                                      cudaMemcpyDeviceToHost, h->d2h_stream));
        return MSED_OK;
    };

    // ---- the call is worked off in rounds ---------------------------------------------------------------
    // A round plans every remaining ode_solver call at the predicted depth (the steps run at dt/4^depth, see
    // plan_depth) as fused launches, enqueues them and asks the controller how far they got.  If a group was
    // not committed -- step s did not go as predicted -- step s alone runs through single attempts, whatever it
    // needs, and the next round plans the rest the way step s went.  (Sub-cycling comes in episodes: before,
    // one wrong prediction sent the rest of the call down the single-attempt path.)
    const bool adaptive = method == MSED_ADAPTIVE_EULER;
    int pred = adaptive ? h->pred_depth : 0;
    int regime = adaptive ? h->regime_depth : 0;
    long long fused_before = 0;     // Ctl::fused_launches when the current round began
    long long base = 0;             // ode_solver calls completed before the current round
    long long nfl_total = 0;        // fused launches enqueued by all rounds
    int nrounds = 0;
    bool first_pending = plan && plan->first && single_attempt;
    const int cur_before = h->cur;
    bool seq_mode = false, seq_committed = false;
    long long seq_nfl = 0;
    bool last_is_fused = false, last_round_clean = false;   // of the latest round
    int depth0 = 0;                 // round 0, for the chunked export's validity test
    long long fused_planned0 = 0;
    bool last_is_fused0 = false;
    const long long max_batch = 256;
    int guard = 0;
    bool stopped = false;
    // what the previous round asks of the next one: plan only that many steps (the good front of a chain that was
    // not committed), or take one step through single attempts (the step behind that front)
    long long next_limit = -1;
    bool next_single = false;
    std::vector<FusedLaunch> fl;
    for (;; ++nrounds) {
    const long long limit = next_limit;
    const bool force_single = next_single;
    next_limit = -1;
    next_single = false;
    const long long rem_all = nsteps - base;
    const long long rem = limit >= 0 ? std::min(limit, rem_all) : rem_all;   // steps this round plans
    int depth = pred;
    double dt_acc = dt;
    long long nq = 1;
    // (Runge-Kutta calls have no accept decision to plan: on a tile small enough for a warp per column they run as
    //  chains of TU_RK_CHAIN_MAX_STEPS calls with the four stages on the column in registers, rk_chain_kernel)
    const bool rk_chain = fusable && !single_attempt && chain_fit && !diag && rem >= 1;
    bool planned = rk_chain || (fusable && single_attempt && !diag && rem >= 1 && !force_single &&
                                plan_depth(dt, h->cfg.dt_min, depth, dt_acc, nq));
    const bool use_chain = planned && chain_fit;
    const int own_rejectable = (adaptive && dt_acc > h->cfg.dt_min) ? 1 : 0;
    fl.clear();
    long long fused_planned = 0;  // steps the fused launches of this round hold
    auto base_commit = [&](long long gate) {
        PlanCommit pc;
        std::memset(&pc, 0, sizeof(pc));
        pc.gate_steps = base + gate;
        pc.own_rejectable = own_rejectable;
        pc.flip = 1;
        pc.launches = 1;
        pc.depth = depth;
        pc.dt_int = 0.0;
        pc.dt_red = dt;
        return pc;
    };
    if (use_chain) {
        // chains cover every step of the round, whatever its parity; a launch holds at most
        // TU_CHAIN_MAX_STEPS accepted sub-steps
        const long long per = rk_chain ? TU_RK_CHAIN_MAX_STEPS : std::max<long long>(1, TU_CHAIN_MAX_STEPS / nq);
        const long long nl = (rem + per - 1) / per;
        const long long each = rem / nl, extra = rem % nl;   // the first `extra` chains hold one step more
        long long gate = 0;
        for (long long q = 0; q < nl; ++q) {
            FusedLaunch f;
            f.m = (int)(each + (q < extra ? 1 : 0));
            f.step = (int)gate;
            f.pc = base_commit(gate);
            f.pc.steps = f.m;
            f.pc.rhs_evals = rk_chain ? 4LL * f.m : f.m * (nq + depth);
            f.pc.subcycles = (long long)f.m * depth;
            f.pc.up_slots = f.m * depth;
            fl.push_back(f);
            gate += f.m;
        }
        fused_planned = rem;
    } else if (planned && depth == 0 && rem >= 2) {
        for (long long q = 0; q < rem / 2; ++q) {
            FusedLaunch f;
            f.kind = PAIR_FULL;
            f.step = (int)(2 * q);
            f.pc = base_commit(2 * q);
            f.pc.steps = 2;
            f.pc.rhs_evals = 2;
            fl.push_back(f);
        }
        fused_planned = 2 * (rem / 2);
    } else if (planned && depth > 0) {
        // every step: the pair that holds the planned rejections and the first two sub-steps, inner pairs,
        // the pair that ends the call (4^depth sub-steps = 4^depth/2 pairs)
        for (long long st = 0; st < rem; ++st) {
            double di = 0.0;
            for (long long j = 0; j < nq / 2; ++j) {
                FusedLaunch f;
                f.kind = (j == 0) ? PAIR_FIRST : (j == nq / 2 - 1 ? PAIR_LAST : PAIR_MID);
                f.step = (int)st;
                f.pc = base_commit(st);
                di = di + dt_acc;            // :138, twice
                di = di + dt_acc;
                f.pc.rhs_evals = 2 + (j == 0 ? depth : 0);
                f.pc.subcycles = (j == 0) ? depth : 0;
                f.pc.up_slots = (j == 0) ? depth : 0;
                if (f.kind == PAIR_LAST) {
                    f.pc.steps = 1;
                } else {
                    f.pc.dt_int = di;
                    f.pc.dt_red = dt_acc;
                }
                fl.push_back(f);
            }
        }
        fused_planned = rem;
    } else {
        planned = false;
    }
    const long long nfl = (long long)fl.size();
    nfl_total += nfl;
    // the call ends with a fused launch, which then leaves the "state of the last get_rhs call" diagnostic
    // behind (KParams::denit_out)
    last_is_fused = nfl > 0 && fused_planned == rem_all;
    if (last_is_fused && (rc = ensure_denit(h))) return rc;
    p.dt_acc = dt_acc;
    p.depth = depth;
    if (nrounds == 0) { depth0 = depth; fused_planned0 = fused_planned; last_is_fused0 = last_is_fused; }

    // Chunk-major Run (msed_run_exchange): when the whole coupling interval is pairs, chunk c runs its boundary
    // assembly, ALL its pairs and its export as soon as its import fields have landed, so that only the first
    // chunk's H2D and the last chunk's D2H are exposed (step-major order leaves the GPU short of work while the
    // transfers of the first pair are still arriving).  The pairs of a chunk go A -> B -> S -> B -> ... through
    // the two state buffers and the staging buffer; the committed state A is untouched until one controller has
    // seen the flags of every pair of every chunk, so a rejected step anywhere still costs nothing but the redo.
    const bool seq_round = nrounds == 0 && plan && plan->stage_private && plan->first && plan->last && first_pending &&
                           !use_chain && nfl >= 2 && last_is_fused && h->scratch != nullptr &&
                           nsteps * depth <= MAX_UP_SLOTS;
    if (seq_round) {
        seq_mode = true;
        seq_nfl = nfl;
        double *A = h->buf[cur_before], *B = h->buf[1 - cur_before], *S = h->scratch;
        // the chunks are independent until the commit: they alternate between two streams, so that the SMs one
        // chunk's last wave leaves idle are taken by the next chunk's first wave (a chunk launch is only a few
        // waves long: 2.4 on C3, where the rounding-up to whole waves cost a quarter of the Run's kernel time)
        const bool two = plan->nchunks >= 2 && h->stream2 != nullptr;
        if (two) {
            CUDA_TRY(h, cudaEventRecord(h->ev_fork, h->stream));
            CUDA_TRY(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
        }
        for (int c = 0; c < plan->nchunks; ++c) {
            cudaStream_t st = (two && (c & 1)) ? h->stream2 : h->stream;
            if ((rc = boundary_chunk(c, st))) return rc;
            const double *in = A;
            for (long long q = 0; q < nfl; ++q) {
                double *out = (q % 2 == 0) ? B : S;
                KParams pc = p;
                pc.col0 = plan->c0[c];
                pc.col_end = plan->c1[c];
                pc.in_ovr = in;
                pc.out_ovr = out;
                pc.pair_kind = fl[q].kind;
                pc.gate_steps = 0;                       // nothing is committed before the whole interval
                pc.up_slot = fl[q].step * depth;
                if (q == nfl - 1) pc.denit_out = h->denit;
                CUDA_TRY(h, launch_pair(h, method, pc, st));
                launches += 1;
                in = out;
            }
            if ((rc = export_chunk(c, st))) return rc;
        }
        if (two) {
            CUDA_TRY(h, cudaEventRecord(h->ev_join, h->stream2));
            CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        }
        first_pending = false;
        if (collective)
            if ((rc = reduce_flags(h, depth > 0 ? MSED_NFLAGS : 8))) return rc;
        // commits all pairs at once (or none)
        PlanCommit pc = base_commit(0);
        pc.steps = nsteps;
        pc.rhs_evals = nsteps * (nq + depth);
        pc.subcycles = nsteps * depth;
        pc.up_slots = (int)(nsteps * depth);
        pc.launches = (int)nfl;
        plan_controller_kernel<<<1, 1, 0, h->stream>>>(h->ctl, pc);
        launches += 1;
    }
    for (long long q = 0; q < (seq_round ? 0 : nfl); ++q) {
        const bool last_launch = last_is_fused && q == nfl - 1;
        const bool chunk_last = nrounds == 0 && last_launch && plan && plan->last;
        KParams pq = p;
        pq.pair_kind = fl[q].kind;
        pq.gate_steps = fl[q].pc.gate_steps;
        pq.up_slot = 0;                                  // every launch is committed on its own: slots start at 0
        if (last_launch) pq.denit_out = h->denit;
        const int m = fl[q].m;
        if (first_pending || chunk_last) {
            for (int c = 0; c < plan->nchunks; ++c) {
                if (first_pending)
                    if ((rc = boundary_chunk(c))) return rc;
                KParams pc = pq;
                pc.col0 = plan->c0[c];
                pc.col_end = plan->c1[c];
                CUDA_TRY(h, rk_chain ? launch_rk_chain(h, method, pc, m, wrapper_clip)
                                     : use_chain ? launch_chain(h, method, pc, m, wrapper_clip) : launch_pair(h, method, pc));
                launches += 1;
                if (chunk_last)
                    if ((rc = export_chunk(c))) return rc;
            }
            first_pending = false;
        } else {
            CUDA_TRY(h, rk_chain ? launch_rk_chain(h, method, pq, m, wrapper_clip)
                                 : use_chain ? launch_chain(h, method, pq, m, wrapper_clip) : launch_pair(h, method, pq));
            launches += 1;
        }
        if (collective)
            if ((rc = reduce_flags(h, fl[q].pc.up_slots > 0 ? MSED_NFLAGS : 8))) return rc;
        plan_controller_kernel<<<1, 1, 0, h->stream>>>(h->ctl, fl[q].pc);
        launches += 1;
    }
    // ---- single attempts: the step a pair plan leaves over (odd count), the step a failed group stopped at,
    // ---- or everything when nothing can be planned ----------------------------------------------------------
    // without a plan: the whole rest for the methods and configurations the fused kernels do not cover; a few
    // steps, then another look at the prediction, when it is the step history that stands in the way
    const bool replannable = fusable && single_attempt && !diag;
    long long target = nsteps;
    if (!planned && replannable) target = std::min(nsteps, base + (force_single ? 1 : 4));
    const long long singles_planned = planned ? (limit >= 0 ? 0 : rem - fused_planned) : target - base;
    if (target != nsteps) {   // (the device copy still holds nsteps otherwise)
        h->ctl_host->steps_target = target;
        CUDA_TRY(h, cudaMemcpyAsync(&h->ctl->steps_target, &h->ctl_host->steps_target, sizeof(long long),
                                    cudaMemcpyHostToDevice, h->stream));
    }
    if (nrounds == 0) CUDA_TRY(h, cudaEventRecord(h->ev_mid, h->stream));

    long long remaining = singles_planned, issued = 0;
    // with nothing but fused launches planned the controller still has to be asked whether all of them were
    // committed; a failed group leaves its step to the single-attempt loop
    bool check_fused = (singles_planned == 0 && nfl > 0);
    bool failure_seen = false;
    // attempts a step is expected to need (sub-cycling: rejected + accepted ones); launches past the end of the
    // call return at once, so a generous estimate only costs empty launches
    long long per_step = 1;
    while (remaining > 0 || check_fused) {
        check_fused = false;
        const long long batch = std::min(remaining * per_step, max_batch);
        for (long long s = 0; s < batch; ++s, ++issued) {
            const bool chunk_first = first_pending && issued == 0;
            const bool chunk_last = nrounds == 0 && plan && plan->last && issued == singles_planned - 1;
            if ((chunk_first || chunk_last) && single_attempt) {
                for (int c = 0; c < plan->nchunks; ++c) {
                    const int c0 = plan->c0[c], c1 = plan->c1[c];
                    if (chunk_first)
                        if ((rc = boundary_chunk(c))) return rc;
                    KParams pc = p;
                    pc.col0 = c0;
                    pc.col_end = c1;
                    CUDA_TRY(h, launch_column(h, method == MSED_EULER ? OP_EULER : OP_ADAPTIVE, pc));
                    launches += 1;
                    if (chunk_last)
                        if ((rc = export_chunk(c))) return rc;
                }
                first_pending = false;
            } else if (method == MSED_EULER) {
                CUDA_TRY(h, launch_column(h, OP_EULER, p));
                launches += 1;
            } else if (method == MSED_ADAPTIVE_EULER) {
                CUDA_TRY(h, launch_column(h, OP_ADAPTIVE, p));
                launches += 1;
            } else if (rk_fused && h->rk_stages == 4 && h->K >= TU_RK_QUAD_MIN_LAYERS) {
                CUDA_TRY(h, launch_rk_quad(h, method, p));   // the whole call in one pass (msed_rkquad.cuh)
                launches += 1;
            } else if (rk_fused) {
                CUDA_TRY(h, launch_rk_pair(h, method, 0, p));
                CUDA_TRY(h, launch_rk_pair(h, method, 1, p));
                launches += 2;
            } else {
                const int first = (method == MSED_RUNGE_KUTTA_4) ? OP_RK4_S1 : OP_RK38_S1;
                for (int st = 0; st < 4; ++st) CUDA_TRY(h, launch_column(h, first + st, p));
                launches += 4;
            }
            if (collective && (method == MSED_ADAPTIVE_EULER || wrapper))
                if ((rc = reduce_flags(h))) return rc;
            controller_kernel<<<1, 1, 0, h->stream>>>(h->ctl, method);
            launches += 1;
            if (diag) {
                minloc_kernel<<<1, 256, 0, h->stream>>>(h->ctl, h->buf[0], h->buf[1], h->ld, h->ncol,
                                                        NV * h->K, h->minloc_val, h->minloc_idx);
                launches += 1;
                if (wrapper) {
                    clip_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->ctl, h->buf[0], h->buf[1],
                                                                         h->mask, h->ld, h->ncol, h->K,
                                                                         minimum);
                    launches += 1;
                }
            }
        }
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaMemcpyAsync(h->ctl_host, h->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        if (h->ctl_host->stop) { stopped = true; break; }
        if (nfl > 0 && h->ctl_host->pairs_disabled && !failure_seen) {
            // a group of this round was not committed: the step it stopped at runs alone through single
            // attempts, the rest of the call is planned afresh afterwards (identical on every rank: the flags
            // the controller decides on are reduced)
            failure_seen = true;
            // how far single attempts go before the rest is planned afresh: the steps the launch that was not
            // committed stood for (a pair: its two steps; a chain: all of its steps -- it cannot be committed in
            // part); a chunk-major interval is all-or-nothing, so it is simply planned again as ordinary pairs,
            // which commit one by one up to the step in question
            long long extent = 1;
            const int fstep = h->ctl_host->fail_step;
            if (seq_round) {
                extent = 0;
            } else if (use_chain && fstep >= 1) {
                // a chain is committed whole or not at all, but it says where it went wrong: the steps in front
                // are run again as a shorter chain (next round), the step in question singly (the round after)
                extent = 0;
                next_limit = fstep;
            } else if (use_chain && fstep == 0) {
                extent = 1;   // its first step: that one singly, then a new plan
            } else
                for (const FusedLaunch &f : fl)
                    if (f.pc.gate_steps == h->ctl_host->steps_done) { extent = std::max<long long>(1, f.pc.steps); break; }
            target = std::min(nsteps, (long long)h->ctl_host->steps_done + extent);
            if (target != h->ctl_host->steps_target) {
                h->ctl_host->steps_target = target;
                CUDA_TRY(h, cudaMemcpyAsync(&h->ctl->steps_target, &h->ctl_host->steps_target, sizeof(long long),
                                            cudaMemcpyHostToDevice, h->stream));
            }
        }
        remaining = target - h->ctl_host->steps_done;
        // sub-cycling needs more attempts than steps: keep going until the controller reports done, and size
        // the next batch by what the steps have needed so far
        if (method == MSED_ADAPTIVE_EULER && remaining > 0) {   // (identical on every rank: the flags are reduced)
            const long long d = h->ctl_host->steps_done > 0 ? h->ctl_host->last_depth : h->ctl_host->step_rej_first;
            per_step = std::max<long long>(per_step, std::min<long long>(d + (1LL << (2 * std::min<long long>(d, 3))), 70));
        }
        if (++guard > 1000000) return fail(h, MSED_ERR_STATE, "step loop did not terminate");
    }
    last_round_clean = !failure_seen && !stopped;
    if (h->ctl_host->fused_launches > fused_before) regime = depth;   // some group of this round was committed
    fused_before = h->ctl_host->fused_launches;
    if (limit >= 0 && !failure_seen) next_single = true;   // the front went through: now the step behind it
    if (nrounds == 0 && seq_mode) seq_committed = last_round_clean && h->ctl_host->steps_done == nsteps;
    if (stopped || nsteps == 0) { ++nrounds; break; }
    base = h->ctl_host->steps_done;
    if (base >= nsteps) { ++nrounds; break; }
    // next round: planned the way the last completed step went; fused launches are allowed again
    // a step that rejected an attempt after its first accepted sub-step says nothing about its successors: back
    // to the regime the last committed group ran in
    if (adaptive) pred = h->ctl_host->last_irregular ? regime : h->ctl_host->last_depth;
    h->ctl_host->pairs_disabled = 0;
    h->ctl_host->steps_target = nsteps;
    CUDA_TRY(h, cudaMemcpyAsync(h->ctl, h->ctl_host, sizeof(Ctl), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));   // ctl_host is read back into by the next round
    if (++guard > 1000000) return fail(h, MSED_ERR_STATE, "step loop did not terminate");
    }  // rounds
    const long long nfl = nfl_total;
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    CUDA_TRY(h, cudaEventSynchronize(h->ev1));
    float ms = 0.f, ms_pairs = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    CUDA_TRY(h, cudaEventElapsedTime(&ms_pairs, h->ev0, h->ev_mid));
    if (nsteps == 0) {
        CUDA_TRY(h, cudaMemcpyAsync(h->ctl_host, h->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }

    const Ctl &r = *h->ctl_host;
    if (plan && plan->last) {  // the chunked export is final only if every attempt went as planned
        CUDA_TRY(h, cudaStreamSynchronize(h->d2h_stream));
        // ... i.e. the call took one round and the launch the export rode on produced the final state: the last
        // fused launch of a committed plan, or the single attempt planned behind it if that was accepted at once
        plan->export_done = single_attempt && !r.stop && r.steps_done == nsteps && nsteps > 0 && r.pair_failures == 0 &&
                            nrounds == 1 && (last_is_fused0 || r.subcycles == fused_planned0 * depth0);
    }
    // the next call is planned the way this call's last step went
    if (r.steps_done > 0 && method == MSED_ADAPTIVE_EULER) {
        h->pred_depth = r.last_irregular ? regime : r.last_depth;
        h->regime_depth = regime;
    }
    h->pairs_committed += r.fused_launches;
    // (the call's last step sits in a committed fused launch of the last round)
    h->denit_valid = last_is_fused && last_round_clean && !r.stop && r.steps_done == nsteps;
    // a committed chunk-major sequence with an even number of pairs ends in the staging buffer: it becomes
    // the state buffer the controller's flipped `cur` points at, the intermediate buffer becomes staging
    if (seq_mode && seq_committed && !r.stop && r.steps_done == nsteps && seq_nfl % 2 == 0) {
        std::swap(h->buf[1 - cur_before], h->scratch);
        // the stepping kernels never write land columns, so the buffer rotated in holds whatever last used the
        // staging area there: restore conc = missing_value (driver :464) before anything can read it
        if (h->has_land) {
            fill_masked_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->buf[1 - cur_before], h->mask, h->ld,
                                                                        h->ncol, NV * h->K, 1.0e20);
            CUDA_TRY(h, cudaGetLastError());
            launches += 1;
        }
    }
    h->cur = r.cur;
    if (diag && r.last_min_dt < h->last_min_dt) {
        // rows are (n*K+k), columns i + inum*j -> Fortran (i,j,k,n), 1-based, in the GLOBAL grid (the tile's
        // origin is cfg.i_offset, cfg.j_offset)
        long long row = -1, gi = 0, gj = 0;
        if (h->comm) {
            // one domain cut into tiles: the minloc of :133 is over all of them.  The condition above is the
            // same on every tile (dt_red and last_min_dt are global), so all tiles reach these reductions.
            // Index = row*2^40 + global column: Fortran array order of the global array for j-slab tiles.
            NcclApi &api = nccl_api();
            double *dval = h->red, *dmin = h->red + 1;
            long long *didx = reinterpret_cast<long long *>(h->red + 2);
            const long long stride = 1LL << 40;
            minloc_pack_kernel<<<1, 1, 0, h->stream>>>(dval, didx, h->minloc_val, h->minloc_idx, h->ncol, stride,
                                                        (long long)h->cfg.j_offset * h->cfg.inum + h->cfg.i_offset);
            int nrc = api.AllReduce(dval, dmin, 1, kNcclFloat64, kNcclMin, h->comm, h->stream);
            minloc_select_kernel<<<1, 1, 0, h->stream>>>(didx, dval, dmin);
            if (nrc == 0) nrc = api.AllReduce(didx, didx, 1, kNcclInt64, kNcclMin, h->comm, h->stream);
            if (nrc != 0) return fail(h, MSED_ERR_NCCL, "ncclAllReduce (minloc) failed");
            long long g = -1;
            CUDA_TRY(h, cudaMemcpyAsync(&g, didx, sizeof(g), cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            if (g != 0x7fffffffffffffffLL && g >= 0) {
                row = g / stride;
                const long long gcol = g % stride;
                gi = gcol % h->cfg.inum;
                gj = gcol / h->cfg.inum;
            }
        } else {
            long long idx = -1;
            CUDA_TRY(h, cudaMemcpy(&idx, h->minloc_idx, sizeof(idx), cudaMemcpyDeviceToHost));
            if (idx >= 0) {
                row = idx / h->ncol;
                const long long col = idx % h->ncol;
                gi = col % h->cfg.inum + h->cfg.i_offset;
                gj = col / h->cfg.inum + h->cfg.j_offset;
            }
        }
        if (row >= 0) {
            h->last_min_dt_grid_cell[0] = (int)gi + 1;
            h->last_min_dt_grid_cell[1] = (int)gj + 1;
            h->last_min_dt_grid_cell[2] = (int)(row % h->K) + 1;
            h->last_min_dt_grid_cell[3] = (int)(row / h->K) + 1;
        }
    }
    h->last_min_dt = r.last_min_dt;
    if (info) {
        info->steps_done = r.steps_done;
        info->rhs_evaluations = r.rhs_evals;
        info->subcycle_warnings = r.subcycles;
        info->last_min_dt = h->last_min_dt;
        for (int q = 0; q < 4; ++q) info->last_min_dt_grid_cell[q] = h->last_min_dt_grid_cell[q];
        info->nan_detected = r.nan_detected;
        info->kernel_ms = ms;
        info->kernel_launches = launches;
        info->fused_pairs = r.fused_launches;
        info->fused_steps = r.fused_steps;
        info->fused_ms = nfl > 0 ? ms_pairs : 0.0;
    }
    if (h->export_pending && !h->export_submitted && !plan) {
        // the next helping of a pending state export: what the copy engine moves in half of the time this call's
        // kernels took (device time, not wall time: a call that waited for copies must not order more of them), so
        // that the next call's own device-to-host copies find the queue empty when they matter.  (A Run with host
        // buffers does this itself, after it has waited for its flux copies: the completion of a stream whose last
        // operation was a copy is signalled through the copy engine's queue, behind whatever was queued before.)
        const int prc = export_pump(h, export_helping(h, (double)ms));
        if (prc) return prc;
    }
    return r.nan_detected ? MSED_NAN_DETECTED : MSED_OK;
}

}  // namespace

// ---- C ABI --------------------------------------------------------------------------------------
extern "C" {

const char *msed_version(void) { return "msed_b200 abi1 sm_100a"; }

size_t msed_sizeof(int what)
{
    return what == 0 ? sizeof(msed_config) : what == 1 ? sizeof(msed_step_info) : 0;
}

const char *msed_last_error(const msed_handle *h) { return h ? h->err.c_str() : g_err.c_str(); }

int msed_config_defaults(msed_config *c)
{
    if (!c) return MSED_ERR_ARG;
    std::memset(c, 0, sizeof(*c));
    c->abi_version = MSED_ABI_VERSION;
    c->inum = c->jnum = 1;
    c->knum = 10;                   // main.F90:49
    c->device = -1;
    c->model = MSED_MODEL_OMEXDIA_P;
    c->dzmin = 0.005;               // main.F90:50
    c->bioturbation_profile = 1;    // fabm_sediment_driver.F90:217-231
    c->diffusivity = 0.9;
    c->bioturbation = 0.9;
    c->bioturbation_depth = 5.0;
    c->bioturbation_min = 0.2;
    c->porosity_max = 0.7;
    c->porosity_fac = 0.9;
    c->k_par = 2.0e-3;
    c->pom_flux_max = 2.0e4;
    c->bioturb_k_l = 0.11;
    c->bioturb_L1 = 0.2;
    c->bioturb_L2 = 0.6;
    c->bioturb_beta = 0.22;
    c->bioturb_b = 1.334;
    c->bioturb_dry_density = 1000.;
    c->distributed_pom_flux = 0;
    c->dt_min = 1.0e-8;             // fabm_sediment_component.F90:60
    c->relative_change_min = -0.9;
    c->bcup_dissolved_variables = 2;  // :64
    c->adaptive_solver_diagnostics = 0;
    c->rLabile = 0.043;             // fabm_sed.nml:51-77
    c->rSemilabile = 0.001;
    c->NCrLdet = 0.22;
    c->NCrSdet = 0.005;
    c->PAds = 0.01;
    c->PAdsODU = 70.;
    c->NH3Ads = 0.0;
    c->CprodMax = 9600.0;
    c->rnit = 200.;
    c->ksO2nitri = 20.;
    c->rODUox = 20.;
    c->ksO2oduox = 1.;
    c->ksO2oxic = 3.;
    c->ksNO3denit = 1.;
    c->kinO2denit = 70.;
    c->kinNO3anox = 1.;
    c->kinO2anox = 1.;
    const double init[NV] = {4.e3, 4.e3, 4.e1, 10., 20., 40., 100., 100.};
    for (int n = 0; n < NV; ++n) { c->initial_value[n] = init[n]; c->minimum[n] = 0.0; }
    return MSED_OK;
}

int msed_create(const msed_config *cfg, msed_handle **out)
{
    if (!cfg || !out) return fail(nullptr, MSED_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->abi_version != MSED_ABI_VERSION) return fail(nullptr, MSED_ERR_ARG, "abi_version mismatch");
    if (cfg->inum < 1 || cfg->jnum < 1) return fail(nullptr, MSED_ERR_ARG, "grid size < 1");  // driver :139
    if (cfg->knum < 2 || cfg->knum > MSED_MAX_LAYERS)
        return fail(nullptr, MSED_ERR_ARG, "knum must be in [2, MSED_MAX_LAYERS]");
    if ((long long)cfg->inum * cfg->jnum > 0x7fffffffLL / 2)
        return fail(nullptr, MSED_ERR_ARG, "tile too large (inum*jnum)");
    if (cfg->model < 0 || cfg->model > 2) return fail(nullptr, MSED_ERR_ARG, "unknown model");
    for (int n = 0; n < NV; ++n)   // the clip compares bit patterns (clip_min, msed_column.cuh); concentrations have no negative floor
        if (!(cfg->minimum[n] >= 0.0)) return fail(nullptr, MSED_ERR_ARG, "state variable minimum must be >= 0");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, MSED_ERR_CUDA, "no CUDA device: libmsed_b200 has no CPU path");
    }
    msed_handle *h = new (std::nothrow) msed_handle();
    if (!h) return fail(nullptr, MSED_ERR_ALLOC, "host allocation failed");
    h->cfg = *cfg;
    for (int n = 0; n < NV; ++n) h->cfg.minimum[n] += 0.0;   // -0.0 -> +0.0
    // auto fusion mode: tiles up to this many columns take the warp-per-column chain kernel (a fifth of it above 32 layers)
    // measured on B200 (profiles/r01_chain_kernel.md): chains win below ~60k columns, where the thread-per-
    // column pair kernel cannot fill the machine (one wave = 148 SMs x 3 CTAs x 128 columns), and lose 5-10 % above
    h->chain_max_cols = 65536;
    if (const char *e = std::getenv("MSED_CHAIN_MAX_COLS")) h->chain_max_cols = std::atoll(e);
    if (const char *e = std::getenv("MSED_EXCHANGE_CHUNK_MAJOR")) h->chunk_major = std::atoi(e) != 0;
    if (const char *e = std::getenv("MSED_RK_STAGES")) h->rk_stages = std::atoi(e) == 2 ? 2 : 4;
    if (const char *e = std::getenv("MSED_EXCHANGE_CHUNKS")) {   // initial msed_set_exchange_chunks value (0 = auto)
        const int n = std::atoi(e);
        if (n >= 0 && n <= 16) h->exchange_chunks = n;
    }
    if (const char *e = std::getenv("MSED_STEP_FUSION")) {  // initial msed_set_step_fusion mode (0..3)
        const int mode = std::atoi(e);
        if (mode >= 0 && mode <= 3) h->step_fusion = mode;
    }
    int dev = cfg->device;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
    h->device = dev;
    h->K = cfg->knum;
    h->ncol = cfg->inum * cfg->jnum;
    h->ld = ((size_t)h->ncol + 15) / 16 * 16;  // 128-byte aligned planes
    h->bioturbation_eff = cfg->bioturbation;
    const int K = h->K;

    // init_grid, driver :147-168
    h->zi.assign(K + 1, 0.0); h->zc.assign(K, 0.0); h->dz.assign(K, 0.0); h->dzc.assign(K - 1, 0.0);
    const double self_fac = 0.18 / ((K + 1) / 2.0 * cfg->dzmin) - 1.0;
    for (int k = 1; k <= K; ++k) {
        h->dz[k - 1] = (1.0 + (self_fac - 1.0) * (double)(k - 1) / (double)(K - 1)) * cfg->dzmin;
        h->zc[k - 1] = h->zi[k - 1] + 0.5 * h->dz[k - 1];
        h->zi[k] = h->zi[k - 1] + h->dz[k - 1];
    }
    for (int k = 0; k < K - 1; ++k) h->dzc[k] = h->zc[k + 1] - h->zc[k];
    // initialize, driver :278-304
    h->bf.assign(K, 1.0); h->por_profile.assign(K, 0.0); h->cumdepth.assign(K, 0.0);
    for (int k = 0; k < K; ++k) {
        h->por_profile[k] = cfg->porosity_max * (1.0 - cfg->porosity_fac * h->zc[k]);
        if (cfg->bioturbation_profile == 1)
            h->bf[k] = std::fmax(cfg->bioturbation_min / cfg->bioturbation,
                                 std::fmax(cfg->bioturbation_depth - 100.0 * h->zi[k], 0.0) /
                                     cfg->bioturbation_depth);
        else if (cfg->bioturbation_profile == 2)
            h->bf[k] = std::exp(-100.0 * h->zi[k] / cfg->bioturbation_depth);
        double s = 0.0;  // sum(dz(:,:,1:k-1)), driver :597
        for (int m = 0; m < k; ++m) s += h->dz[m];
        h->cumdepth[k] = s;
    }

#define CREATE_TRY(expr)                                                                          \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            int rc_ = fail(nullptr, MSED_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(e_)); \
            msed_destroy(h);                                                                      \
            return rc_;                                                                           \
        }                                                                                         \
    } while (0)

    CREATE_TRY(cudaSetDevice(h->device));
    CREATE_TRY(enable_pair_smem());
    CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    CREATE_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    if (!(std::getenv("MSED_EXCHANGE_TWO_STREAMS") && std::atoi(std::getenv("MSED_EXCHANGE_TWO_STREAMS")) == 0))
        CREATE_TRY(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    for (auto &e : h->ev_pool) CREATE_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : h->ev_x) CREATE_TRY(cudaEventCreate(&e));
    CREATE_TRY(cudaEventCreate(&h->ev0));
    CREATE_TRY(cudaEventCreate(&h->ev_mid));
    CREATE_TRY(cudaEventCreate(&h->ev1));
    const size_t state_bytes = (size_t)NV * K * h->ld * sizeof(double);
    CREATE_TRY(cudaMalloc(&h->buf[0], state_bytes));
    CREATE_TRY(cudaMalloc(&h->buf[1], state_bytes));
    CREATE_TRY(cudaMalloc(&h->por, (size_t)K * h->ld * sizeof(double)));
    CREATE_TRY(cudaMalloc(&h->bdys, (size_t)(NV + 1) * h->ld * sizeof(double)));
    CREATE_TRY(cudaMalloc(&h->fluxes, (size_t)NV * h->ld * sizeof(double)));
    CREATE_TRY(cudaMalloc(&h->par_surface, h->ld * sizeof(double)));
    CREATE_TRY(cudaMalloc(&h->mask, h->ld));
    CREATE_TRY(cudaMalloc(&h->tables, (size_t)3 * MAXK * sizeof(double)));
    CREATE_TRY(cudaMalloc(&h->ctl, sizeof(Ctl)));
    CREATE_TRY(cudaMalloc(&h->minloc_val, sizeof(double)));
    CREATE_TRY(cudaMalloc(&h->minloc_idx, sizeof(long long)));
    CREATE_TRY(cudaMalloc(&h->red, 32 * sizeof(double)));
    CREATE_TRY(cudaMallocHost(&h->ctl_host, sizeof(Ctl)));
    CREATE_TRY(cudaMemsetAsync(h->buf[0], 0, state_bytes, h->stream));  // conc = 0.0_rk, component :531
    CREATE_TRY(cudaMemsetAsync(h->buf[1], 0, state_bytes, h->stream));
    CREATE_TRY(cudaMemsetAsync(h->bdys, 0, (size_t)(NV + 1) * h->ld * sizeof(double), h->stream));
    CREATE_TRY(cudaMemsetAsync(h->fluxes, 0, (size_t)NV * h->ld * sizeof(double), h->stream));
    CREATE_TRY(cudaMemsetAsync(h->par_surface, 0, h->ld * sizeof(double), h->stream));
    CREATE_TRY(cudaMemsetAsync(h->mask, 0, h->ld, h->stream));
    CREATE_TRY(cudaMemsetAsync(h->ctl, 0, sizeof(Ctl), h->stream));
    CREATE_TRY(cudaMemsetAsync(h->minloc_idx, 0xff, sizeof(long long), h->stream));
    std::vector<double> t(3 * MAXK, 0.0);
    for (int k = 0; k < K; ++k) { t[k] = h->zc[k]; t[MAXK + k] = h->cumdepth[k]; t[2 * MAXK + k] = h->por_profile[k]; }
    CREATE_TRY(cudaMemcpyAsync(h->tables, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    fill_porosity_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->por, h->mask, h->ld, h->ncol, K,
                                                                  h->tables + 2 * MAXK);
    CREATE_TRY(cudaGetLastError());
    CREATE_TRY(cudaStreamSynchronize(h->stream));
#undef CREATE_TRY
    *out = h;
    return MSED_OK;
}

int msed_destroy(msed_handle *h)
{
    if (!h) return MSED_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm) { nccl_api().CommDestroy(h->comm); h->comm = nullptr; }
    cudaFree(h->buf[0]); cudaFree(h->buf[1]); cudaFree(h->aux[0]); cudaFree(h->aux[1]);
    cudaFree(h->por); cudaFree(h->bdys); cudaFree(h->fluxes); cudaFree(h->par_surface);
    cudaFree(h->scratch); cudaFree(h->denit); cudaFree(h->tables); cudaFree(h->mask); cudaFree(h->colmap); cudaFree(h->xstage); cudaFree(h->ctl);
    cudaFree(h->minloc_val); cudaFree(h->minloc_idx); cudaFree(h->red); cudaFree(h->pel); cudaFree(h->snap);
    if (h->ev_snap) cudaEventDestroy(h->ev_snap);
    if (h->ev_export) cudaEventDestroy(h->ev_export);
    if (h->export_stream) cudaStreamDestroy(h->export_stream);
    if (h->ctl_host) cudaFreeHost(h->ctl_host);
    for (auto &e : h->ev_pool) if (e) cudaEventDestroy(e);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev_mid) cudaEventDestroy(h->ev_mid);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MSED_OK;
}

int msed_get_grid(const msed_handle *h, double *zi, double *zc, double *dz, double *dzc)
{
    if (!h) return MSED_ERR_ARG;
    if (zi) std::memcpy(zi, h->zi.data(), h->zi.size() * sizeof(double));
    if (zc) std::memcpy(zc, h->zc.data(), h->zc.size() * sizeof(double));
    if (dz) std::memcpy(dz, h->dz.data(), h->dz.size() * sizeof(double));
    if (dzc) std::memcpy(dzc, h->dzc.data(), h->dzc.size() * sizeof(double));
    return MSED_OK;
}

int msed_set_mask(msed_handle *h, const int32_t *mask2d)
{
    if (!h || !mask2d) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    std::vector<unsigned char> m(h->ld, 0);
    for (int c = 0; c < h->ncol; ++c) m[c] = mask2d[c] > 0 ? 1 : 0;
    CUDA_TRY(h, cudaMemcpyAsync(h->mask, m.data(), h->ld, cudaMemcpyHostToDevice, h->stream));
    apply_mask_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->por, h->buf[0], h->buf[1], h->mask, h->ld,
                                                               h->ncol, h->K);
    CUDA_TRY(h, cudaGetLastError());
    // list of the wet columns for pair_kernel (see there); a tile without land keeps the identity
    h->wet_idx.clear();
    for (int c = 0; c < h->ncol; ++c)
        if (!m[c]) h->wet_idx.push_back(c);
    h->has_land = (int)h->wet_idx.size() < h->ncol;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->colmap) { cudaFree(h->colmap); h->colmap = nullptr; }
    if ((int)h->wet_idx.size() < h->ncol && !h->wet_idx.empty() && !std::getenv("MSED_NO_COLMAP")) {
        CUDA_TRY(h, cudaMalloc(&h->colmap, h->wet_idx.size() * sizeof(int)));
        CUDA_TRY(h, cudaMemcpy(h->colmap, h->wet_idx.data(), h->wet_idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    return MSED_OK;
}

int msed_set_porosity(msed_handle *h, const double *porosity3d)
{
    if (!h || !porosity3d) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = upload_rows(h, h->por, porosity3d, h->K);
    if (rc) return rc;
    h->por_mode = 0;  // arbitrary field: stream it
    apply_mask_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->por, h->buf[0], h->buf[1], h->mask, h->ld,
                                                               h->ncol, h->K);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_update_porosity_from_surface(msed_handle *h, const double *porosity_surface2d)
{
    if (!h || !porosity_surface2d) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->scratch, porosity_surface2d, (size_t)h->ncol * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    porosity_from_surface_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(
        h->por, h->scratch, h->mask, h->ld, h->ncol, h->K, h->cfg.porosity_fac, h->tables);
    h->por_mode = 2;  // porosity(:,:,k) = porosity(:,:,1) * (1 - porosity_fac*(zc(k)-zc(1)))
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_set_par_surface(msed_handle *h, const double *par_surface2d)
{
    if (!h || !par_surface2d) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(h->par_surface, par_surface2d, (size_t)h->ncol * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_check_domain(msed_handle *h)
{
    if (!h) return MSED_ERR_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    // grid conditions (driver :513-530) are properties of init_grid's closed form
    for (int k = 0; k < h->K - 1; ++k)
        if (h->dzc[k] <= 0) return fail(h, MSED_BAD_DOMAIN, "sediment central layer difference <= 0");
    for (int k = 0; k < h->K; ++k)
        if (h->dz[k] < h->cfg.dzmin) return fail(h, MSED_BAD_DOMAIN, "sediment layer height < minimum value");
    // porosity conditions (driver :503-511) are checked where the field lives
    int *dflags = (int *)h->minloc_idx;  // 8 bytes of device scratch: [0] porosity<=0, [1] porosity>1
    CUDA_TRY(h, cudaMemsetAsync(dflags, 0, 2 * sizeof(int), h->stream));
    check_porosity_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->por, h->mask, h->ld, h->ncol, h->K, dflags);
    CUDA_TRY(h, cudaGetLastError());
    int hflags[2] = {0, 0};
    CUDA_TRY(h, cudaMemcpyAsync(hflags, dflags, sizeof(hflags), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->minloc_idx, 0xff, sizeof(long long), h->stream));
    if (hflags[0]) return fail(h, MSED_BAD_DOMAIN, "sediment porosity <=0");
    if (hflags[1]) return fail(h, MSED_BAD_DOMAIN, "sediment porosity > 1");
    apply_mask_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->por, h->buf[0], h->buf[1], h->mask, h->ld,
                                                               h->ncol, h->K);  // driver :532-541
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_init_concentrations(msed_handle *h)
{
    if (!h) return MSED_ERR_ARG;
    h->denit_valid = false;
    CUDA_TRY(h, cudaSetDevice(h->device));
    InitVals iv;
    for (int n = 0; n < NV; ++n) iv.v[n] = h->cfg.initial_value[n];
    init_conc_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->buf[0], h->buf[1], h->por, h->mask, h->ld,
                                                              h->ncol, h->K, iv, 1.e20);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_set_state(msed_handle *h, const double *conc)
{
    if (!h || !conc) return fail(h, MSED_ERR_ARG, "null argument");
    h->denit_valid = false;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = upload_rows(h, h->buf[h->cur], conc, (size_t)NV * h->K);
    if (rc) return rc;
    // masked columns must hold the same values in both buffers (the step kernels skip them)
    copy_state_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->buf[1 - h->cur], h->buf[h->cur], h->ld,
                                                               h->ncol, NV * h->K);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_get_state(msed_handle *h, double *conc)
{
    if (!h || !conc) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    return download_rows(h, conc, h->buf[h->cur], (size_t)NV * h->K);
}

int msed_set_state_from_column(msed_handle *h, const double *conc1d)
{
    if (!h || !conc1d) return fail(h, MSED_ERR_ARG, "null argument");
    h->denit_valid = false;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->scratch, conc1d, (size_t)NV * h->K * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    broadcast_column_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->buf[0], h->buf[1], h->scratch,
                                                                     h->mask, h->ld, h->ncol, h->K);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_set_boundary(msed_handle *h, const double *bdys, const double *fluxes)
{
    if (!h) return MSED_ERR_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc;
    if (bdys && (rc = upload_rows(h, h->bdys, bdys, NV + 1))) return rc;
    if (fluxes && (rc = upload_rows(h, h->fluxes, fluxes, NV))) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_set_boundary_device(msed_handle *h, const void *bdys_dev, const void *fluxes_dev)
{
    if (!h) return MSED_ERR_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t w = (size_t)h->ncol * sizeof(double);
    if (bdys_dev)
        CUDA_TRY(h, cudaMemcpy2DAsync(h->bdys, h->ld * sizeof(double), bdys_dev, w, w, NV + 1,
                                      cudaMemcpyDeviceToDevice, h->stream));
    if (fluxes_dev)
        CUDA_TRY(h, cudaMemcpy2DAsync(h->fluxes, h->ld * sizeof(double), fluxes_dev, w, w, NV,
                                      cudaMemcpyDeviceToDevice, h->stream));
    return MSED_OK;
}

int msed_get_fluxes_device(msed_handle *h, void *fluxes_dev)
{
    if (!h || !fluxes_dev) return MSED_ERR_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t w = (size_t)h->ncol * sizeof(double);
    CUDA_TRY(h, cudaMemcpy2DAsync(fluxes_dev, w, h->fluxes, h->ld * sizeof(double), w, NV,
                                  cudaMemcpyDeviceToDevice, h->stream));
    return MSED_OK;
}

int msed_get_boundary_conditions(msed_handle *h, const double *temperature2d, const double *const *csurf,
                                 const double *const *wz)
{
    if (!h) return MSED_ERR_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h);
    if (rc) return rc;
    // pack the import fields that are present into the staging buffer (dense rows of ncol)
    BcPtrs in;
    std::memset(&in, 0, sizeof(in));
    double *stage = h->scratch;
    size_t row = 0;
    const size_t w = (size_t)h->ncol;
    auto push = [&](const double *src, const double **dst) -> int {
        CUDA_TRY(h, cudaMemcpyAsync(stage + row * w, src, w * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        *dst = stage + row * w;
        ++row;
        return MSED_OK;
    };
    if (temperature2d && (rc = push(temperature2d, &in.temperature))) return rc;
    for (int n = 0; n < NV; ++n) {
        if (!csurf || !csurf[n]) continue;
        if (n < NPART) {
            if (!wz || !wz[n]) return fail(h, MSED_ERR_ARG, "particulate variable without z_velocity field");
            if ((rc = push(wz[n], &in.wz[n]))) return rc;
        }
        if ((rc = push(csurf[n], &in.csurf[n]))) return rc;
    }
    boundary_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(
        h->bdys, h->fluxes, h->buf[h->cur], h->por, in, h->ld, w, h->ncol, h->K,
        h->cfg.bcup_dissolved_variables, h->bioturbation_eff, h->cfg.diffusivity, h->dz[0]);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_get_boundary(msed_handle *h, double *bdys, double *fluxes)
{
    if (!h) return MSED_ERR_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc;
    if (bdys && (rc = download_rows(h, bdys, h->bdys, NV + 1))) return rc;
    if (fluxes && (rc = download_rows(h, fluxes, h->fluxes, NV))) return rc;
    return MSED_OK;
}

int msed_get_fluxes(msed_handle *h, double *fluxes)
{
    if (!h || !fluxes) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    return download_rows(h, fluxes, h->fluxes, NV);
}

int msed_get_upward_fluxes(msed_handle *h, double *upward)
{
    if (!h || !upward) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h);
    if (rc) return rc;
    negate_rows_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->scratch, h->fluxes, h->ld, h->ld, h->ncol, NV);
    CUDA_TRY(h, cudaGetLastError());
    return download_rows(h, upward, h->scratch, NV);
}

int msed_get_field(msed_handle *h, int which, double *out3d)
{
    if (!h || !out3d) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n2 = (size_t)h->ncol;
    if (which == MSED_FIELD_POROSITY) return download_rows(h, out3d, h->por, h->K);
    if (which == MSED_FIELD_LAYER_HEIGHT || which == MSED_FIELD_LAYER_CENTER_DEPTH) {
        const std::vector<double> &t = (which == MSED_FIELD_LAYER_HEIGHT) ? h->dz : h->zc;
        for (int k = 0; k < h->K; ++k)
            for (size_t c = 0; c < n2; ++c) out3d[(size_t)k * n2 + c] = t[k];
        return MSED_OK;
    }
    if (which < 0 || which > MSED_FIELD_FLUX_CAP) return fail(h, MSED_ERR_ARG, "unknown field");
    int rc = ensure_scratch(h);
    if (rc) return rc;
    KParams p;
    fill_params(h, p);
    // FABM diagnostics describe the state of the last get_rhs call: for Euler / adaptive Euler
    // that is the buffer the accepted attempt read; for RK the last stage state c1 (same buffer)
    const double *state = h->buf[h->cur];
    if (which == MSED_FIELD_DENIT) state = h->buf[1 - h->cur];
    if (which == MSED_FIELD_DENIT && h->denit_valid) {  // the call ended with a fused pair: stored field
        denit_copy_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->scratch, h->denit, h->mask, h->ld,
                                                                   h->ncol, h->K);
        CUDA_TRY(h, cudaGetLastError());
        return download_rows(h, out3d, h->scratch, h->K);
    }
    field_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->scratch, which, p, h->par_surface, state,
                                                          h->tables + MAXK, h->cfg.k_par, -999.0);
    CUDA_TRY(h, cudaGetLastError());
    return download_rows(h, out3d, h->scratch, h->K);
}

int msed_get_rhs(msed_handle *h, double *rhs)
{
    if (!h || !rhs) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h);
    if (rc) return rc;
    if (h->cfg.bioturbation_profile == 3) h->bioturbation_eff = 1.0;
    KParams p;
    fill_params(h, p);
    p.use_ctl = 0;
    p.buf[0] = h->buf[h->cur];
    p.buf[1] = h->buf[1 - h->cur];
    CUDA_TRY(h, launch_column(h, OP_RHS, p));
    rc = download_rows(h, rhs, h->scratch, (size_t)NV * h->K);
    if (rc) return rc;
    // diagnostics now describe the current state: mirror it into the "last rhs state" buffer (and drop the
    // diagnostic a fused launch may have left behind for the state before)
    h->denit_valid = false;
    copy_state_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(h->buf[1 - h->cur], h->buf[h->cur], h->ld,
                                                               h->ncol, NV * h->K);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_ode_solver(msed_handle *h, double dt, int method, msed_step_info *info)
{
    if (!h) return MSED_ERR_ARG;
    return run_steps(h, dt, method, 1, false, info);
}

int msed_step(msed_handle *h, double dt, int method, int64_t nsteps, msed_step_info *info)
{
    if (!h) return MSED_ERR_ARG;
    return run_steps(h, dt, method, nsteps, true, info);
}

int msed_run(msed_handle *h, double dt, int method, double run_seconds, msed_step_info *info)
{
    if (!h) return MSED_ERR_ARG;
    if (!(dt > 0.0) || run_seconds < 0.0) return fail(h, MSED_ERR_ARG, "dt <= 0 or run_seconds < 0");
    // component :1700-1769: full steps of dt, the last one shortened to hit stopTime (:1705-1708)
    long long nfull = (long long)std::floor(run_seconds / dt * (1.0 + 1e-14));
    double rem = run_seconds - (double)nfull * dt;
    if (rem < 1e-9 * dt) rem = 0.0;
    msed_step_info a, b;
    std::memset(&a, 0, sizeof(a));
    std::memset(&b, 0, sizeof(b));
    int rc = run_steps(h, dt, method, nfull, true, &a);
    if (rc == MSED_OK && rem > 0.0) rc = run_steps(h, rem, method, 1, true, &b);
    if (info) {
        *info = a;
        if (rem > 0.0) {
            info->steps_done += b.steps_done;
            info->rhs_evaluations += b.rhs_evaluations;
            info->subcycle_warnings += b.subcycle_warnings;
            info->last_min_dt = b.last_min_dt;
            for (int q = 0; q < 4; ++q) info->last_min_dt_grid_cell[q] = b.last_min_dt_grid_cell[q];
            info->nan_detected |= b.nan_detected;
            info->kernel_ms += b.kernel_ms;
            info->kernel_launches += b.kernel_launches;
            info->fused_pairs += b.fused_pairs;
            info->fused_steps += b.fused_steps;
            info->fused_ms += b.fused_ms;
        }
    }
    return rc;
}

int msed_set_step_fusion(msed_handle *h, int enable)
{
    if (!h) return MSED_ERR_ARG;
    if (enable < 0 || enable > 3) return fail(h, MSED_ERR_ARG, "step fusion mode must be 0..3");
    h->step_fusion = enable;
    h->pred_depth = 0;
    h->regime_depth = 0;
    return MSED_OK;
}

int msed_set_rk_stages_per_launch(msed_handle *h, int stages)
{
    if (!h) return MSED_ERR_ARG;
    if (stages != 2 && stages != 4) return fail(h, MSED_ERR_ARG, "Runge-Kutta stages per launch must be 2 or 4");
    h->rk_stages = stages;
    return MSED_OK;
}

int msed_set_exchange_order(msed_handle *h, int chunk_major)
{
    if (!h) return MSED_ERR_ARG;
    h->chunk_major = chunk_major ? 1 : 0;
    return MSED_OK;
}

int msed_set_exchange_chunks(msed_handle *h, int nchunks)
{
    if (!h || nchunks < 0 || nchunks > 16) return fail(h, MSED_ERR_ARG, "nchunks must be in [0,16]");
    h->exchange_chunks = nchunks;
    return MSED_OK;
}

int msed_run_exchange(msed_handle *h, double dt, int method, double run_seconds, const double *temperature2d,
                      const double *const *csurf, const double *const *wz, double *upward_fluxes,
                      msed_step_info *info)
{
    if (!h || !upward_fluxes) return fail(h, MSED_ERR_ARG, "null argument");
    if (!(dt > 0.0) || run_seconds < 0.0) return fail(h, MSED_ERR_ARG, "dt <= 0 or run_seconds < 0");
    CUDA_TRY(h, cudaSetDevice(h->device));
    long long nfull = (long long)std::floor(run_seconds / dt * (1.0 + 1e-14));
    double rem = run_seconds - (double)nfull * dt;
    if (rem < 1e-9 * dt) rem = 0.0;
    int nchunks = h->exchange_chunks;
    // chosen from the tile size: small tiles are one chunk -- nothing to overlap, but the asynchronous
    // sequence below still saves the host synchronisations of the three separate calls, which dominate a
    // Run of a few tens of microseconds; msed_set_exchange_chunks(1) asks for the plain sequence
    const bool auto_chunks = nchunks == 0;
    // (C3, 10^6 columns: e2e 52.5 G cell-updates/s with 6 or 8 chunks, 50.0 with 4, 47.5 with 3 -- profiles/r02_s24_*)
    if (auto_chunks) nchunks = h->ncol >= (1 << 19) ? 8 : (h->ncol >= (1 << 17) ? 4 : 1);
    const bool pipelined = (nchunks > 1 || auto_chunks) && (method == MSED_EULER || method == MSED_ADAPTIVE_EULER) &&
                           !h->cfg.adaptive_solver_diagnostics && (nfull + (rem > 0.0 ? 1 : 0)) > 0 &&
                           (size_t)NV * h->K >= 12 + NV;
    int rc;
    if (!pipelined) {  // plain sequence: boundary, loop, export
        if ((rc = msed_get_boundary_conditions(h, temperature2d, csurf, wz))) return rc;
        rc = msed_run(h, dt, method, run_seconds, info);
        if (rc < 0) return rc;
        const int rc2 = msed_get_upward_fluxes(h, upward_fluxes);
        return rc2 ? rc2 : rc;
    }
    if ((rc = ensure_scratch(h))) return rc;
    ExchangePlan plan;
    std::memset(&plan.bc, 0, sizeof(plan.bc));
    plan.nchunks = nchunks;
    // Chunk boundaries: equal column counts.  (MSED_EXCHANGE_WAVE_CHUNKS=1 cuts the chunks at whole waves of
    // pair_kernel CTAs -- SMs x resident CTAs x columns per CTA -- in the index space the launches run over, the
    // wet-column list on a tile with land, so that only the last chunk ends in a partly filled wave.  Measured: C3
    // e2e 46.8 against 50.6 G cell-updates/s, C4 slab 80.9 against 80.4 -- the two compute streams of the chunk-major
    // Run already fill one chunk's last wave with the next chunk's first, and the larger first chunk starts later.)
    int used = 0;
    {
        static const bool wave_off = !(std::getenv("MSED_EXCHANGE_WAVE_CHUNKS") && std::atoi(std::getenv("MSED_EXCHANGE_WAVE_CHUNKS")) != 0);
        if (wave_off) {
            const int per = ((h->ncol + nchunks - 1) / nchunks + COL_BLOCK - 1) / COL_BLOCK * COL_BLOCK;
            for (int c = 0; c < nchunks; ++c) {
                const int c0 = c * per, c1 = std::min(h->ncol, (c + 1) * per);
                if (c0 >= c1) break;
                plan.c0[used] = c0;
                plan.c1[used] = c1;
                ++used;
            }
        } else {
            const bool by_wet = h->colmap != nullptr && !h->wet_idx.empty();
            const long long work = by_wet ? (long long)h->wet_idx.size() : (long long)h->ncol;
            int nsm = 148;
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
            const long long wave = (long long)nsm * TU_PAIR_CTAS_PER_SM * COL_BLOCK;
            long long per = (work + nchunks - 1) / nchunks;
            const long long unit = per >= wave ? wave : COL_BLOCK;   // (a tile of less than a wave per chunk: as before)
            per = (per + unit - 1) / unit * unit;
            for (int c = 0; c < nchunks; ++c) {
                const long long w0 = (long long)c * per, w1 = std::min(work, (long long)(c + 1) * per);
                if (w0 >= w1) break;
                // column range of the chunk: from its first unit of work to the next chunk's (land in front of the
                // first wet column goes with the first chunk, land behind the last one with the last)
                plan.c0[used] = (c == 0) ? 0 : (by_wet ? h->wet_idx[(size_t)w0] : (int)w0);
                plan.c1[used] = (w1 >= work) ? h->ncol : (by_wet ? h->wet_idx[(size_t)w1] : (int)w1);
                ++used;
            }
        }
    }
    plan.nchunks = used;
    // staging rows: [0..11] import fields, [12..19] negated fluxes
    double *stage = h->scratch;
    if (h->chunk_major) {  // the staging buffer proper is a state buffer of the chunk-major sequence (run_steps)
        if (!h->xstage) CUDA_TRY(h, cudaMalloc(&h->xstage, (size_t)(12 + NV) * h->ld * sizeof(double)));
        stage = h->xstage;
        plan.stage_private = true;
    }
    const double *host[12];
    const double **dev[12];
    int key[12];   // index of the field in the generation table (msed_set_import_generations)
    int nf = 0;
    if (temperature2d) { host[nf] = temperature2d; dev[nf] = &plan.bc.temperature; key[nf] = 0; ++nf; }
    for (int n = 0; n < NV; ++n) {
        if (!csurf || !csurf[n]) continue;
        if (n < NPART) {
            if (!wz || !wz[n]) return fail(h, MSED_ERR_ARG, "particulate variable without z_velocity field");
            host[nf] = wz[n]; dev[nf] = &plan.bc.wz[n]; key[nf] = 2 + 2 * n; ++nf;
        }
        host[nf] = csurf[n]; dev[nf] = &plan.bc.csurf[n]; key[nf] = 1 + 2 * n; ++nf;
    }
    for (int f = 0; f < nf; ++f) *dev[f] = stage + (size_t)f * h->ld;
    // A tile of one chunk is a Run of a few hundred microseconds, of which a dozen small copies and their staging
    // are a third: when every field the caller handed over is pinned host memory, the boundary kernel reads the
    // import fields, and the export kernel writes the fluxes, directly through PCIe instead.
    {
        static const bool off = std::getenv("MSED_EXCHANGE_ZERO_COPY") && std::atoi(std::getenv("MSED_EXCHANGE_ZERO_COPY")) == 0;
        bool zc = plan.nchunks == 1 && !off;
        const void *devp[12];
        void *outp = nullptr;
        for (int f = 0; f <= nf && zc; ++f) {
            cudaPointerAttributes at;
            const void *q = f < nf ? (const void *)host[f] : (const void *)upward_fluxes;
            if (cudaPointerGetAttributes(&at, q) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
                cudaGetLastError();
                zc = false;
            } else if (f < nf) {
                devp[f] = at.devicePointer;
            } else {
                outp = at.devicePointer;
            }
        }
        if (zc) {
            for (int f = 0; f < nf; ++f) *dev[f] = (const double *)devp[f];
            plan.zero_copy = true;
            plan.host_out_dev = (double *)outp;
        }
    }
    // a field the caller has not changed since its last upload (same counter, same host array, same staging
    // row of a staging area nothing else writes) is already on the device
    bool resident[12];
    for (int f = 0; f < nf; ++f) {
        msed_handle::ImportRow &r = h->import_row[key[f]];
        resident[f] = plan.zero_copy || (h->import_gen_on && plan.stage_private && r.row == f && r.stage == stage &&
                                         r.host == host[f] && r.gen == h->import_gen[key[f]]);
        r.host = host[f];
        r.gen = h->import_gen[key[f]];
        r.row = (h->import_gen_on && plan.stage_private && !plan.zero_copy) ? f : -1;
        r.stage = stage;
    }
    plan.neg = stage + (size_t)12 * h->ld;
    plan.host_out = upward_fluxes;
    CUDA_TRY(h, cudaEventRecord(h->ev_x[0], h->copy_stream));
    for (int c = 0; c < plan.nchunks; ++c) {  // all H2D traffic on the copy stream, chunk by chunk
        const int c0 = plan.c0[c], n = plan.c1[c] - plan.c0[c];
        for (int f = 0; f < nf; ++f)
            if (!resident[f])
                CUDA_TRY(h, cudaMemcpyAsync(stage + (size_t)f * h->ld + c0, host[f] + c0, (size_t)n * sizeof(double),
                                            cudaMemcpyHostToDevice, h->copy_stream));
        CUDA_TRY(h, cudaEventRecord(h->ev_pool[c], h->copy_stream));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_x[1], h->copy_stream));
    msed_step_info a, b;
    std::memset(&a, 0, sizeof(a));
    std::memset(&b, 0, sizeof(b));
    bool exported = false;
    rc = MSED_OK;
    if (nfull > 0) {
        plan.first = true;
        plan.last = (rem == 0.0);
        rc = run_steps(h, dt, method, nfull, true, &a, &plan);
        exported = plan.export_done;
    }
    if (rc == MSED_OK && rem > 0.0) {
        plan.first = (nfull == 0);
        plan.last = true;
        plan.export_done = false;
        rc = run_steps(h, rem, method, 1, true, &b, &plan);
        exported = plan.export_done;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_x[2], h->d2h_stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->d2h_stream));
    {   // phase marks of this Run (msed_get_exchange_timing); ev0 / ev1 bracket the kernels of the last run_steps
        float t01 = 0.f, t0k = 0.f, tk2 = 0.f, t02 = 0.f;
        cudaEventElapsedTime(&t01, h->ev_x[0], h->ev_x[1]);
        cudaEventElapsedTime(&t0k, h->ev_x[0], h->ev1);
        cudaEventElapsedTime(&tk2, h->ev1, h->ev_x[2]);
        cudaEventElapsedTime(&t02, h->ev_x[0], h->ev_x[2]);
        cudaGetLastError();
        h->exchange_ms[0] = t01; h->exchange_ms[1] = t0k; h->exchange_ms[2] = tk2; h->exchange_ms[3] = t02;
    }
    if (rc < 0) return rc;
    if (!exported) {  // an attempt was rejected (or NaN): export again from the final state
        const int rc2 = msed_get_upward_fluxes(h, upward_fluxes);
        if (rc2) return rc2;
    }
    if (h->export_pending && !h->export_submitted) {   // see run_steps
        const int prc = export_pump(h, export_helping(h, a.kernel_ms + b.kernel_ms));
        if (prc) return prc;
    }
    if (info) {
        *info = a;
        if (rem > 0.0) {
            info->steps_done += b.steps_done;
            info->rhs_evaluations += b.rhs_evaluations;
            info->subcycle_warnings += b.subcycle_warnings;
            info->last_min_dt = b.last_min_dt;
            for (int q = 0; q < 4; ++q) info->last_min_dt_grid_cell[q] = b.last_min_dt_grid_cell[q];
            info->nan_detected |= b.nan_detected;
            info->kernel_ms += b.kernel_ms;
            info->kernel_launches += b.kernel_launches;
            info->fused_pairs += b.fused_pairs;
            info->fused_steps += b.fused_steps;
            info->fused_ms += b.fused_ms;
        }
    }
    return rc;
}

int msed_get_exchange_timing(const msed_handle *h, double *ms4)
{
    if (!h || !ms4) return MSED_ERR_ARG;
    for (int q = 0; q < 4; ++q) ms4[q] = h->exchange_ms[q];
    return MSED_OK;
}

int msed_set_import_generations(msed_handle *h, const uint64_t *gen)
{
    if (!h) return fail(h, MSED_ERR_ARG, "null handle");
    h->import_gen_on = gen != nullptr;
    for (int q = 0; q < 1 + 2 * NV; ++q) h->import_gen[q] = gen ? (unsigned long long)gen[q] : 0ULL;
    if (!gen)
        for (auto &r : h->import_row) r = msed_handle::ImportRow();
    return MSED_OK;
}

int msed_spinup_column(const msed_config *cfg, const double *bdys1d, const double *fluxes1d, int64_t nsteps,
                       int method, double *conc1d, msed_step_info *info)
{
    if (!cfg || !bdys1d || !fluxes1d || !conc1d) return fail(nullptr, MSED_ERR_ARG, "null argument");
    // One column is a batch of one: the whole pre-simulation in a single launch with the column in registers
    // (2 years: about 1 s) instead of a launch per attempt on a 1x1 tile (40 s; bit-identical,
    // tests/test_gpu_spinup.py).  MSED_SPINUP_PATH=steps keeps the launch-per-attempt path, which also serves the
    // configurations outside the batch kernel's scope.
    {
        const char *e = std::getenv("MSED_SPINUP_PATH");
        const bool steps = e && !std::strcmp(e, "steps");
        if (!steps && cfg->knum <= TU_SPINUP_MAX_LAYERS && !cfg->distributed_pom_flux &&
            cfg->model != MSED_MODEL_TEST_SOLVER)
            return msed_spinup_batch(cfg, 1, nullptr, bdys1d, fluxes1d, nsteps, method, conc1d, info);
    }
    // sed1d: a 1x1xknum clone with Dirichlet boundaries, constant bioturbation and solver
    // diagnostics switched on (component :557-611); dt_spinup = 3600 s (:574)
    msed_config c1 = *cfg;
    c1.inum = c1.jnum = 1;
    c1.i_offset = c1.j_offset = 0;
    c1.bcup_dissolved_variables = 2;
    c1.adaptive_solver_diagnostics = 1;
    msed_handle *h = nullptr;
    int rc = msed_create(&c1, &h);
    if (rc) return rc;
    // sed1d%bioturbation_profile=0 is set AFTER initialize (:611): the bioturbation_factor table
    // keeps the shape it was initialised with; only the profile-3 branch is switched off
    h->cfg.bioturbation_profile = (c1.bioturbation_profile == 3) ? 0 : c1.bioturbation_profile;
    if ((rc = msed_check_domain(h)) || (rc = msed_init_concentrations(h)) ||
        (rc = msed_set_boundary(h, bdys1d, fluxes1d))) {
        g_err = h->err;
        msed_destroy(h);
        return rc;
    }
    rc = run_steps(h, 3600.0, method, nsteps, false, info);  // ode_solver only, no clipping (:614-618)
    if (rc == MSED_OK) rc = msed_get_state(h, conc1d);
    if (rc) g_err = h->err;
    msed_destroy(h);
    return rc;
}

namespace {
void apply_member(msed_config &c, const msed_spinup_member &m)
{
    c.rLabile = m.rLabile; c.rSemilabile = m.rSemilabile; c.NCrLdet = m.NCrLdet; c.NCrSdet = m.NCrSdet;
    c.PAds = m.PAds; c.PAdsODU = m.PAdsODU; c.NH3Ads = m.NH3Ads; c.CprodMax = m.CprodMax; c.rnit = m.rnit;
    c.ksO2nitri = m.ksO2nitri; c.rODUox = m.rODUox; c.ksO2oduox = m.ksO2oduox; c.ksO2oxic = m.ksO2oxic;
    c.ksNO3denit = m.ksNO3denit; c.kinO2denit = m.kinO2denit; c.kinNO3anox = m.kinNO3anox; c.kinO2anox = m.kinO2anox;
    for (int n = 0; n < NV; ++n) c.initial_value[n] = m.initial_value[n];
}
}  // namespace

int msed_spinup_batch(const msed_config *cfg, int32_t nmembers, const msed_spinup_member *members,
                      const double *bdys1d, const double *fluxes1d, int64_t nsteps, int method, double *conc1d,
                      msed_step_info *info)
{
    if (!cfg || !bdys1d || !fluxes1d || !conc1d || nmembers < 1 || nsteps < 0)
        return fail(nullptr, MSED_ERR_ARG, "bad arguments");
    if (method < 0 || method > 3) return fail(nullptr, MSED_ERR_ARG, "unknown ode_method");
    const int K = cfg->knum;
    const size_t P = (size_t)nmembers;
    // configurations outside the batch kernel's scope: member by member through msed_spinup_column
    if (K > TU_SPINUP_MAX_LAYERS || cfg->distributed_pom_flux || cfg->model == MSED_MODEL_TEST_SOLVER) {
        std::vector<double> bd(NV + 1), fl(NV), col((size_t)NV * K);
        for (size_t m = 0; m < P; ++m) {
            msed_config cm = *cfg;
            if (members) apply_member(cm, members[m]);
            for (int n = 0; n <= NV; ++n) bd[n] = bdys1d[m + P * n];
            for (int n = 0; n < NV; ++n) fl[n] = fluxes1d[m + P * n];
            const int rc = msed_spinup_column(&cm, bd.data(), fl.data(), nsteps, method, col.data(), info ? &info[m] : nullptr);
            if (rc) return rc;
            for (size_t q = 0; q < (size_t)NV * K; ++q) conc1d[m + P * q] = col[q];
        }
        return MSED_OK;
    }
    // sed1d: 1x1xknum clones with Dirichlet boundaries, constant bioturbation and solver diagnostics
    // (component :557-611), laid out as one nmembers x 1 tile
    msed_config c1 = *cfg;
    c1.inum = nmembers;
    c1.jnum = 1;
    c1.i_offset = c1.j_offset = 0;
    c1.bcup_dissolved_variables = 2;
    c1.adaptive_solver_diagnostics = 1;
    msed_handle *h = nullptr;
    int rc = msed_create(&c1, &h);
    if (rc) return rc;
    h->cfg.bioturbation_profile = (c1.bioturbation_profile == 3) ? 0 : c1.bioturbation_profile;   // :611
    OmexDev *d_om = nullptr;
    double *d_lmd = nullptr;
    int *d_cell = nullptr;
    long long *d_cnt = nullptr;
    auto cleanup = [&](int code) {
        if (code) g_err = h->err.empty() ? g_err : h->err;
        cudaFree(d_om); cudaFree(d_lmd); cudaFree(d_cell); cudaFree(d_cnt);
        msed_destroy(h);
        return code;
    };
#define BATCH_TRY(expr)                                                                                   \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess) return cleanup(fail(h, MSED_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(e_))); \
    } while (0)
    if ((rc = msed_check_domain(h)) || (rc = msed_init_concentrations(h)) || (rc = msed_set_boundary(h, bdys1d, fluxes1d)))
        return cleanup(rc);
    if (members) {   // per-member reaction parameters and initial values (init_concentrations, driver :456)
        std::vector<OmexDev> om(P);
        std::vector<double> conc((size_t)NV * K * P);
        for (size_t m = 0; m < P; ++m) {
            msed_config cm = *cfg;
            apply_member(cm, members[m]);
            om[m] = omex_dev(cm);
            for (int n = 0; n < NV; ++n)
                for (int k = 0; k < K; ++k) conc[m + P * (k + (size_t)K * n)] = cm.initial_value[n] / h->por_profile[k];
        }
        BATCH_TRY(cudaMalloc(&d_om, P * sizeof(OmexDev)));
        BATCH_TRY(cudaMemcpy(d_om, om.data(), P * sizeof(OmexDev), cudaMemcpyHostToDevice));
        if ((rc = msed_set_state(h, conc.data()))) return cleanup(rc);
    }
    BATCH_TRY(cudaMalloc(&d_lmd, P * sizeof(double)));
    BATCH_TRY(cudaMalloc(&d_cell, 4 * P * sizeof(int)));
    BATCH_TRY(cudaMalloc(&d_cnt, 2 * P * sizeof(long long)));
    KParams p;
    fill_params(h, p);
    p.use_ctl = 0;
    p.buf[0] = h->buf[h->cur];
    p.buf[1] = h->buf[1 - h->cur];
    SpinupArgs a;
    a.om = d_om;
    a.last_min_dt = d_lmd;
    a.grid_cell = d_cell;
    a.counters = d_cnt;
    a.nsteps = nsteps;
    a.dt = 3600.0;                        // dt_spinup, :574
    a.dt_min = h->cfg.dt_min;
    a.last_min_dt0 = (double)1.e20f;      // solver_library.F90:44
    a.method = method;
    BATCH_TRY(cudaEventRecord(h->ev0, h->stream));
    BATCH_TRY(tu_launch_spinup(h->cfg.model, p, a, h->stream));
    BATCH_TRY(cudaEventRecord(h->ev1, h->stream));
    BATCH_TRY(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    BATCH_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if ((rc = msed_get_state(h, conc1d))) return cleanup(rc);
    if (info) {
        std::vector<double> lmd(P);
        std::vector<int> cell(4 * P);
        std::vector<long long> cnt(2 * P);
        BATCH_TRY(cudaMemcpy(lmd.data(), d_lmd, P * sizeof(double), cudaMemcpyDeviceToHost));
        BATCH_TRY(cudaMemcpy(cell.data(), d_cell, 4 * P * sizeof(int), cudaMemcpyDeviceToHost));
        BATCH_TRY(cudaMemcpy(cnt.data(), d_cnt, 2 * P * sizeof(long long), cudaMemcpyDeviceToHost));
        for (size_t m = 0; m < P; ++m) {
            std::memset(&info[m], 0, sizeof(msed_step_info));
            info[m].steps_done = nsteps;
            info[m].rhs_evaluations = cnt[2 * m];
            info[m].subcycle_warnings = cnt[2 * m + 1];
            info[m].last_min_dt = lmd[m];
            for (int q = 0; q < 4; ++q) info[m].last_min_dt_grid_cell[q] = cell[4 * m + q];
            info[m].kernel_ms = ms;
            info[m].kernel_launches = 1;
        }
    }
#undef BATCH_TRY
    return cleanup(MSED_OK);
}

int msed_pelagic_init(msed_handle *h, const double *conc2d, const double *wz2d, const double *layer_height2d,
                      const double *temperature2d)
{
    if (!h || !conc2d || !wz2d || !layer_height2d || !temperature2d) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t rows = 2 * NV + 2;
    if (!h->pel) {
        CUDA_TRY(h, cudaMalloc(&h->pel, rows * h->ld * sizeof(double)));
        CUDA_TRY(h, cudaMemsetAsync(h->pel, 0, rows * h->ld * sizeof(double), h->stream));
    }
    int rc;
    if ((rc = upload_rows(h, h->pel, conc2d, NV))) return rc;
    if ((rc = upload_rows(h, h->pel + (size_t)NV * h->ld, wz2d, NV))) return rc;
    if ((rc = upload_rows(h, h->pel + (size_t)2 * NV * h->ld, layer_height2d, 1))) return rc;
    if ((rc = upload_rows(h, h->pel + (size_t)(2 * NV + 1) * h->ld, temperature2d, 1))) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_pelagic_get(msed_handle *h, double *conc2d)
{
    if (!h || !conc2d) return fail(h, MSED_ERR_ARG, "null argument");
    if (!h->pel) return fail(h, MSED_ERR_STATE, "msed_pelagic_init has not been called");
    CUDA_TRY(h, cudaSetDevice(h->device));
    return download_rows(h, conc2d, h->pel, NV);
}

int msed_coupled_run(msed_handle *h, double dt, int method, double coupling_seconds, int64_t ncouplings,
                     msed_step_info *info)
{
    if (!h) return MSED_ERR_ARG;
    if (!h->pel) return fail(h, MSED_ERR_STATE, "msed_pelagic_init has not been called");
    if (!(dt > 0.0) || !(coupling_seconds > 0.0) || ncouplings < 0) return fail(h, MSED_ERR_ARG, "bad arguments");
    CUDA_TRY(h, cudaSetDevice(h->device));
    msed_step_info acc, one;
    std::memset(&acc, 0, sizeof(acc));
    acc.last_min_dt = h->last_min_dt;
    for (int q = 0; q < 4; ++q) acc.last_min_dt_grid_cell[q] = h->last_min_dt_grid_cell[q];
    BcPtrs in;
    std::memset(&in, 0, sizeof(in));
    in.temperature = h->pel + (size_t)(2 * NV + 1) * h->ld;
    for (int n = 0; n < NV; ++n) {
        in.csurf[n] = h->pel + (size_t)n * h->ld;
        if (n < NPART) in.wz[n] = h->pel + (size_t)(NV + n) * h->ld;
    }
    int rc = MSED_OK;
    for (int64_t c = 0; c < ncouplings && rc == MSED_OK; ++c) {
        boundary_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(
            h->bdys, h->fluxes, h->buf[h->cur], h->por, in, h->ld, h->ld, h->ncol, h->K,
            h->cfg.bcup_dissolved_variables, h->bioturbation_eff, h->cfg.diffusivity, h->dz[0]);
        CUDA_TRY(h, cudaGetLastError());
        rc = msed_run(h, dt, method, coupling_seconds, &one);
        if (rc < 0) return rc;
        pelagic_flux_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(
            h->pel, h->fluxes, h->pel + (size_t)2 * NV * h->ld, h->mask, h->ld, h->ncol, coupling_seconds);
        CUDA_TRY(h, cudaGetLastError());
        acc.steps_done += one.steps_done;
        acc.rhs_evaluations += one.rhs_evaluations;
        acc.subcycle_warnings += one.subcycle_warnings;
        acc.last_min_dt = one.last_min_dt;
        for (int q = 0; q < 4; ++q) acc.last_min_dt_grid_cell[q] = one.last_min_dt_grid_cell[q];
        acc.nan_detected |= one.nan_detected;
        acc.kernel_ms += one.kernel_ms;
        acc.kernel_launches += one.kernel_launches + 2;
        acc.fused_pairs += one.fused_pairs;
        acc.fused_steps += one.fused_steps;
        acc.fused_ms += one.fused_ms;
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (info) *info = acc;
    return rc;
}

int msed_pelagic_benthic_coupler(msed_handle *h, const msed_pelagic_state *st)
{
    if (!h || !st) return fail(h, MSED_ERR_ARG, "null argument");
    if (!st->temperature || !st->oxygen || !st->detN || !st->detN_z_velocity)
        return fail(h, MSED_ERR_ARG, "temperature, oxygen, detN and detN_z_velocity are required");
    if ((!st->nitrate || !st->ammonium || !st->DIP) && !st->DIN)
        return fail(h, MSED_ERR_ARG, "DIN is required when nitrate, ammonium or DIP is absent");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h);
    if (rc) return rc;
    // staging rows (row length ld): [0..10] inputs, [11] temperature, [12..19] csurf, [20..22] wz
    const size_t ld = h->ld, w = (size_t)h->ncol;
    double *stage = h->scratch;
    if ((size_t)NV * h->K < 23) return fail(h, MSED_ERR_STATE, "staging buffer too small");
    auto up = [&](const double *src, size_t row, const double **dst) -> int {
        *dst = nullptr;
        if (!src) return MSED_OK;
        CUDA_TRY(h, cudaMemcpyAsync(stage + row * ld, src, w * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        *dst = stage + row * ld;
        return MSED_OK;
    };
    P2BIn in;
    const double *temp = nullptr;
    if ((rc = up(st->oxygen, 0, &in.oxygen)) || (rc = up(st->detN, 1, &in.detN)) ||
        (rc = up(st->detN_z_velocity, 2, &in.detN_wz)) || (rc = up(st->detC, 3, &in.detC)) ||
        (rc = up(st->detP, 4, &in.detP)) || (rc = up(st->detP_z_velocity, 5, &in.detP_wz)) ||
        (rc = up(st->nitrate, 6, &in.nitrate)) || (rc = up(st->ammonium, 7, &in.ammonium)) ||
        (rc = up(st->DIN, 8, &in.DIN)) || (rc = up(st->DIP, 9, &in.DIP)) || (rc = up(st->temperature, 11, &temp)))
        return rc;
    double *csurf = stage + 12 * ld, *wz = stage + 20 * ld;
    pelagic_benthic_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(csurf, wz, in, ld, h->ncol,
                                                                    (h->compat & MSED_COMPAT_P2B_OXYGEN_LAST_CELL) ? 1 : 0);
    CUDA_TRY(h, cudaGetLastError());
    BcPtrs bc;
    std::memset(&bc, 0, sizeof(bc));
    bc.temperature = temp;
    for (int n = 0; n < NV; ++n) {
        bc.csurf[n] = csurf + (size_t)n * ld;
        if (n < NPART) bc.wz[n] = wz + (size_t)n * ld;
    }
    boundary_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(
        h->bdys, h->fluxes, h->buf[h->cur], h->por, bc, h->ld, ld, h->ncol, h->K,
        h->cfg.bcup_dissolved_variables, h->bioturbation_eff, h->cfg.diffusivity, h->dz[0]);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_benthic_pelagic_coupler(msed_handle *h, const msed_benthic_pelagic_params *par,
                                 const msed_pelagic_fluxes *out)
{
    if (!h || !par || !out) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h);
    if (rc) return rc;
    const double dip = par->dipflux_const < 0.0 ? par->dinflux_const / 16.0 : par->dipflux_const;
    benthic_pelagic_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(
        h->scratch, h->fluxes, h->ld, h->ncol, par->dinflux_const, dip, par->convertN, par->NC_fdet,
        par->NC_sdet);
    CUDA_TRY(h, cudaGetLastError());
    double *dst[8] = {out->nitrate, out->ammonium, out->DIN, out->DIP, out->detN, out->detC, out->detP,
                      out->oxygen};
    for (int r = 0; r < 8; ++r)
        if (dst[r])
            CUDA_TRY(h, cudaMemcpyAsync(dst[r], h->scratch + (size_t)r * h->ld, (size_t)h->ncol * sizeof(double),
                                        cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_soil_pelagic_connector(msed_handle *h, const msed_soil_pelagic_params *par,
                                const msed_soil_pelagic_fluxes *out)
{
    if (!h || !par || !out) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h);
    if (rc) return rc;
    // InitializeP1 :156: a negative dipflux_const follows dinflux_const in Redfield proportion
    const double dip = par->dipflux_const < 0.0 ? par->dinflux_const / 16.0 : par->dipflux_const;
    // 86400.0*365.0 is a default-real product in the reference (:469,:530); 31536000 is exact in binary32
    const double year = 86400.0 * 365.0;
    const int oxy_mode = (out->oxygen ? 1 : 0) | (out->odu ? 2 : 0);
    soil_pelagic_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(
        h->scratch, h->fluxes, h->ld, h->ncol, par->dinflux_const / year, dip / year, par->convertN,
        par->convertP, oxy_mode);
    CUDA_TRY(h, cudaGetLastError());
    struct { double *dst; int row; } rows[9] = {
        {out->nitrate, 0}, {out->ammonium, 1}, {out->DIN, 2}, {out->DIP, 3}, {out->oxygen, 4},
        {out->odu, 5},     {out->detC, 6},     {out->detN, 7}, {out->detP, 7}};
    for (int r = 0; r < 9; ++r)
        if (rows[r].dst)
            CUDA_TRY(h, cudaMemcpyAsync(rows[r].dst, h->scratch + (size_t)rows[r].row * h->ld,
                                        (size_t)h->ncol * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_export_state_begin(msed_handle *h, double *conc_host)
{
    if (!h || !conc_host) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->export_pending) {
        const int rc = msed_export_state_wait(h);
        if (rc) return rc;
    }
    if (!h->ev_snap) {
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_snap, cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_export, cudaEventDisableTiming));
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->export_stream, cudaStreamNonBlocking));
    }
    const size_t bytes = (size_t)NV * h->K * h->ld * sizeof(double);
    if (!h->snap && !h->export_direct) {
        // a state-sized snapshot lets the PCIe copy run under the following Runs (it takes several of them);
        // without the memory for it the copy reads the state itself and the next stepping call waits for it
        if (cudaMalloc(&h->snap, bytes) != cudaSuccess) {
            cudaGetLastError();
            h->snap = nullptr;
            h->export_direct = true;
        }
    }
    const double *src = h->buf[h->cur];
    if (h->snap) {
        CUDA_TRY(h, cudaMemcpyAsync(h->snap, src, bytes, cudaMemcpyDeviceToDevice, h->stream));
        src = h->snap;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_snap, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->export_stream, h->ev_snap, 0));
    // in column slices of about 32 MB, a few at a time (export_pump)
    {
        h->export_slice = ((size_t)32 << 20) / sizeof(double);   // elements per copy
        h->export_src = src;
        h->export_dst = conc_host;
        h->export_next = 0;
        h->export_submitted = false;
        h->export_pending = true;
        // without a snapshot the state must stay put until the copy is through: everything at once, and the
        // compute stream waits; with one, a first helping now and the rest under the calls that follow
        const int rc = export_pump(h, h->snap ? 128.0e6 : -1.0);
        if (rc) return rc;
        if (!h->snap) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_export, 0));
    }
    h->export_pending = true;
    return MSED_OK;
}

int msed_export_state_wait(msed_handle *h)
{
    if (!h) return MSED_ERR_ARG;
    if (!h->export_pending) return MSED_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int rc = export_pump(h, -1.0);
    if (rc) return rc;
    CUDA_TRY(h, cudaEventSynchronize(h->ev_export));
    h->export_pending = false;
    return MSED_OK;
}

int msed_set_compat(msed_handle *h, int flags)
{
    if (!h) return MSED_ERR_ARG;
    if (flags & ~(MSED_COMPAT_P2B_OXYGEN_LAST_CELL | MSED_COMPAT_P2S_HEAD)) return fail(h, MSED_ERR_ARG, "unknown compat flag");
    h->compat = flags;
    return MSED_OK;
}

int msed_pelagic_soil_params_defaults(msed_pelagic_soil_params *par)
{
    if (!par) return MSED_ERR_ARG;
    par->sinking_factor = 0.3;                      // pelagic_soil_connector.F90:38
    par->NC_ldet = 0.23;                            // :39
    par->NC_sdet = 0.01;                            // :40
    par->convertN = 1.0;                            // :41
    par->convertP = 1.0;                            // :42
    par->sinking_factor_min = (double)0.02f;        // :43 (default-real literal)
    par->half_sedimentation_depth = (double)0.1f;   // :44
    par->half_sedimentation_tke = 1.0e3;            // :45
    par->critical_detritus = (double)60.0f;         // :46
    return MSED_OK;
}

int msed_pelagic_soil_connector(msed_handle *h, const msed_pelagic_soil_state *st, const msed_pelagic_soil_params *par)
{
    if (!h || !st || !par) return fail(h, MSED_ERR_ARG, "null argument");
    if (!st->temperature || !st->detN || !st->detN_z_velocity)
        return fail(h, MSED_ERR_ARG, "temperature, detN and detN_z_velocity are required");
    if (!st->nitrate && !st->ammonium && !st->DIN)
        return fail(h, MSED_ERR_ARG, "one of nitrate, ammonium, DIN is required");   // ESMF_RC_NOT_FOUND, :1846
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h);
    if (rc) return rc;
    // staging rows (row length ld): [0..14] inputs, [15] temperature, [16..23] csurf, [24..26] wz
    const size_t ld = h->ld, w = (size_t)h->ncol;
    double *stage = h->scratch;
    if ((size_t)NV * h->K < 27) return fail(h, MSED_ERR_STATE, "staging buffer too small");
    auto up = [&](const double *src, size_t row, const double **dst) -> int {
        *dst = nullptr;
        if (!src) return MSED_OK;
        CUDA_TRY(h, cudaMemcpyAsync(stage + row * ld, src, w * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        *dst = stage + row * ld;
        return MSED_OK;
    };
    P2SIn in;
    const double *temp = nullptr, *par2d = nullptr;
    if ((rc = up(st->oxygen, 0, &in.oxygen)) || (rc = up(st->odu, 1, &in.odu)) || (rc = up(st->detN, 2, &in.detN)) ||
        (rc = up(st->detN_z_velocity, 3, &in.detN_wz)) || (rc = up(st->detC, 4, &in.detC)) ||
        (rc = up(st->detP, 5, &in.detP)) || (rc = up(st->detP_z_velocity, 6, &in.detP_wz)) ||
        (rc = up(st->nitrate, 7, &in.nitrate)) || (rc = up(st->ammonium, 8, &in.ammonium)) ||
        (rc = up(st->DIN, 9, &in.DIN)) || (rc = up(st->DIP, 10, &in.DIP)) || (rc = up(st->water_depth, 11, &in.depth)) ||
        (rc = up(st->tke, 12, &in.tke)) || (rc = up(st->par, 13, &par2d)) || (rc = up(st->temperature, 15, &temp)))
        return rc;
    P2SPar q;
    q.sinking_factor = par->sinking_factor;
    q.sinking_factor_min = par->sinking_factor_min;
    q.NC_ldet = par->NC_ldet;
    q.NC_sdet = par->NC_sdet;
    q.half_sedimentation_depth = par->half_sedimentation_depth;
    q.half_sedimentation_tke = par->half_sedimentation_tke;
    q.critical_detritus = par->critical_detritus;
    q.convertN = par->convertN;
    q.convertP = par->convertP;
    q.head_compat = (h->compat & MSED_COMPAT_P2S_HEAD) ? 1 : 0;
    double *csurf = stage + 16 * ld, *wz = stage + 24 * ld;
    if (q.head_compat)   // the carbon velocity fields keep the 0 the component created them with
        CUDA_TRY(h, cudaMemsetAsync(wz, 0, 2 * ld * sizeof(double), h->stream));
    pelagic_soil_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(csurf, wz, nullptr, in, q, ld, h->ncol);
    CUDA_TRY(h, cudaGetLastError());
    if (par2d)
        CUDA_TRY(h, cudaMemcpyAsync(h->par_surface, par2d, w * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    BcPtrs bc;
    std::memset(&bc, 0, sizeof(bc));
    bc.temperature = temp;
    for (int n = 0; n < NV; ++n) {
        // oxygen / reduced substances stay untouched when neither is imported (:862-900 maps a 2-D field only)
        if (n >= 6 && !st->oxygen && !st->odu) continue;
        bc.csurf[n] = csurf + (size_t)n * ld;
        if (n < NPART) bc.wz[n] = wz + (size_t)n * ld;
    }
    boundary_kernel<<<nblocks(h->ncol), 256, 0, h->stream>>>(
        h->bdys, h->fluxes, h->buf[h->cur], h->por, bc, h->ld, ld, h->ncol, h->K,
        h->cfg.bcup_dissolved_variables, h->bioturbation_eff, h->cfg.diffusivity, h->dz[0]);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_diagnostics(msed_handle *h, double *bed_flux_sum, double *inventory, int reduce_over_ranks)
{
    if (!h || (!bed_flux_sum && !inventory)) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h);
    if (rc) return rc;
    KParams p;
    fill_params(h, p);
    const int nb = nblocks(h->ncol);
    diag_partial_kernel<<<nb, 256, 0, h->stream>>>(h->scratch, h->buf[h->cur], h->por, h->fluxes, h->mask, h->ld,
                                                   h->ncol, h->K, p);
    diag_final_kernel<<<1, 32, 0, h->stream>>>(h->red, h->scratch, nb);
    CUDA_TRY(h, cudaGetLastError());
    if (reduce_over_ranks && h->comm) {
        NcclApi &api = nccl_api();
        if (api.AllReduce(h->red, h->red, 2 * NV, kNcclFloat64, kNcclSum, h->comm, h->stream) != 0)
            return fail(h, MSED_ERR_NCCL, "ncclAllReduce (diagnostics) failed");
    }
    double out[2 * NV];
    CUDA_TRY(h, cudaMemcpyAsync(out, h->red, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int n = 0; n < NV; ++n) {
        if (bed_flux_sum) bed_flux_sum[n] = out[n];
        if (inventory) inventory[n] = out[NV + n];
    }
    return MSED_OK;
}

int msed_state_checksum(msed_handle *h, int64_t global_ncol, int64_t col_offset, uint64_t out[2])
{
    if (!h || !out) return fail(h, MSED_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    unsigned long long *d = reinterpret_cast<unsigned long long *>(h->red);
    CUDA_TRY(h, cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), h->stream));
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    checksum_kernel<<<nsm * 8, 256, 0, h->stream>>>(h->buf[h->cur], h->mask, h->ld, h->ncol, NV * h->K,
                                                    (long long)global_ncol, (long long)col_offset, d);
    CUDA_TRY(h, cudaGetLastError());
    unsigned long long r[2];
    CUDA_TRY(h, cudaMemcpyAsync(r, d, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    out[0] = r[0];
    out[1] = r[1];
    return MSED_OK;
}

int msed_measure_fp64_peak(int device, double *tflops)
{
    if (!tflops) return MSED_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, MSED_ERR_CUDA, "no CUDA device");
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    CUDA_TRY(nullptr, cudaSetDevice(device));
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    double *out = nullptr;
    CUDA_TRY(nullptr, cudaMalloc(&out, sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(nullptr, cudaEventCreate(&e0));
    CUDA_TRY(nullptr, cudaEventCreate(&e1));
    const int iters = 1 << 15, threads = 256, blocks = nsm * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {   // the first launch warms up
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, threads>>>(out, 0.999999, 1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma = (double)blocks * threads * 8.0 * iters;
        if (rep > 0 && ms > 0.f) best = std::max(best, 2.0 * fma / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    CUDA_TRY(nullptr, cudaGetLastError());
    *tflops = best;
    return MSED_OK;
}

int msed_set_stream(msed_handle *h, void *cuda_stream)
{
    if (!h) return MSED_ERR_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    h->stream = (cudaStream_t)cuda_stream;
    return MSED_OK;
}

int msed_synchronize(msed_handle *h)
{
    if (!h) return MSED_ERR_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MSED_OK;
}

int msed_device_state(msed_handle *h, void **dev_ptr, size_t *ld)
{
    if (!h || !dev_ptr || !ld) return MSED_ERR_ARG;
    *dev_ptr = h->buf[h->cur];
    *ld = h->ld;
    return MSED_OK;
}

int msed_nccl_unique_id(char id[128])
{
    NcclApi &api = nccl_api();
    if (!api.ok) return fail(nullptr, MSED_ERR_NCCL, "libnccl.so.2 not found");
    NcclUniqueId u;
    int rc = api.GetUniqueId(&u);
    if (rc != 0) return fail(nullptr, MSED_ERR_NCCL, "ncclGetUniqueId failed");
    std::memcpy(id, u.internal, 128);
    return MSED_OK;
}

int msed_comm_init(msed_handle *h, const char id[128], int nranks, int rank)
{
    if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, MSED_ERR_ARG, "bad comm arguments");
    NcclApi &api = nccl_api();
    if (!api.ok) return fail(h, MSED_ERR_NCCL, "libnccl.so.2 not found");
    CUDA_TRY(h, cudaSetDevice(h->device));
    NcclUniqueId u;
    std::memcpy(u.internal, id, 128);
    int rc = api.CommInitRank(&h->comm, nranks, u, rank);
    if (rc != 0) {
        h->comm = nullptr;
        return fail(h, MSED_ERR_NCCL, std::string("ncclCommInitRank: ") +
                                          (api.GetErrorString ? api.GetErrorString(rc) : "error"));
    }
    h->nranks = nranks;
    h->rank = rank;
    return MSED_OK;
}

int msed_comm_destroy(msed_handle *h)
{
    if (!h) return MSED_ERR_ARG;
    if (h->comm) { nccl_api().CommDestroy(h->comm); h->comm = nullptr; }
    h->nranks = 1;
    h->rank = 0;
    return MSED_OK;
}

int msed_set_allreduce_hook(msed_handle *h, msed_allreduce_hook hook, void *user)
{
    if (!h) return MSED_ERR_ARG;
    h->hook = hook;
    h->hook_user = user;
    return MSED_OK;
}

}  // extern "C"
