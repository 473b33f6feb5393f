// msed_tu_chain.cu -- instantiations of chain_kernel (msed_chain.cuh) and their launcher.
#include "msed_launch.h"

namespace msed {
#include "msed_column.cuh"
#include "msed_chain.cuh"

static_assert(TU_CHAIN_MAX_LAYERS == CHAIN_MAX_LAYERS, "msed_launch.h and msed_chain.cuh disagree");
static_assert(TU_CHAIN_MAX_STEPS == MSED_CHAIN_MAX_STEPS, "msed_launch.h and msed_chain.cuh disagree");

cudaError_t tu_launch_chain(int model, bool adaptive, bool clip, const KParams &p, int m, cudaStream_t s)
{
    if (p.col_end <= p.col0) return cudaSuccess;
    const dim3 grid((p.col_end - p.col0 + CHAIN_WARPS - 1) / CHAIN_WARPS), block(CHAIN_BLOCK);
    const bool sub = adaptive && p.depth > 0;
    const bool two = p.K > 32;   // two layers per lane
#define MSED_CHAIN(MODEL, AD, CL, SB)                                                    \
    do {                                                                                 \
        if (two) chain_kernel<MODEL, AD, CL, SB, 2><<<grid, block, 0, s>>>(p, m);        \
        else chain_kernel<MODEL, AD, CL, SB, 1><<<grid, block, 0, s>>>(p, m);            \
    } while (0)
#define MSED_CHAIN_CL(MODEL, AD, CL) do { if (sub) MSED_CHAIN(MODEL, AD, CL, true); else MSED_CHAIN(MODEL, AD, CL, false); } while (0)
#define MSED_CHAIN_MODEL(MODEL)                                                                              \
    do {                                                                                                     \
        if (adaptive) { if (clip) MSED_CHAIN_CL(MODEL, true, true); else MSED_CHAIN_CL(MODEL, true, false); } \
        else          { if (clip) MSED_CHAIN(MODEL, false, true, false); else MSED_CHAIN(MODEL, false, false, false); } \
    } while (0)
    if (model == MSED_MODEL_OMEXDIA_P) MSED_CHAIN_MODEL(MSED_MODEL_OMEXDIA_P);
    else MSED_CHAIN_MODEL(MSED_MODEL_NONE);
#undef MSED_CHAIN_MODEL
#undef MSED_CHAIN_CL
#undef MSED_CHAIN
    return cudaGetLastError();
}

cudaError_t tu_launch_rk_chain(int model, int method, bool clip, const KParams &p, int m, cudaStream_t s)
{
    if (p.col_end <= p.col0) return cudaSuccess;
    const dim3 grid((p.col_end - p.col0 + CHAIN_WARPS - 1) / CHAIN_WARPS), block(CHAIN_BLOCK);
    const bool two = p.K > 32;
#define MSED_RKC(MODEL, METH, CL)                                                          \
    do {                                                                                   \
        if (two) rk_chain_kernel<MODEL, METH, CL, 2><<<grid, block, 0, s>>>(p, m);         \
        else rk_chain_kernel<MODEL, METH, CL, 1><<<grid, block, 0, s>>>(p, m);             \
    } while (0)
#define MSED_RKC_M(MODEL)                                                                  \
    do {                                                                                   \
        if (method == MSED_RUNGE_KUTTA_4) { if (clip) MSED_RKC(MODEL, MSED_RUNGE_KUTTA_4, true); else MSED_RKC(MODEL, MSED_RUNGE_KUTTA_4, false); } \
        else { if (clip) MSED_RKC(MODEL, MSED_RUNGE_KUTTA_4_38, true); else MSED_RKC(MODEL, MSED_RUNGE_KUTTA_4_38, false); } \
    } while (0)
    if (model == MSED_MODEL_OMEXDIA_P) MSED_RKC_M(MSED_MODEL_OMEXDIA_P);
    else MSED_RKC_M(MSED_MODEL_NONE);
#undef MSED_RKC_M
#undef MSED_RKC
    return cudaGetLastError();
}

}  // namespace msed
