// msed_tu_rk.cu -- instantiations of rk_pair_kernel (msed_rkpair.cuh) and their launcher.
#include "msed_launch.h"

namespace msed {
#include "msed_column.cuh"
#include "msed_pair.cuh"   // sts64 and the shared-memory geometry
#include "msed_rkpair.cuh"

cudaError_t tu_launch_rk_pair(int model, int method, int which, const KParams &p, cudaStream_t s)
{
    if (p.col_end <= p.col0) return cudaSuccess;
    const dim3 grid((p.col_end - p.col0 + COL_BLOCK - 1) / COL_BLOCK), block(COL_BLOCK);
    const int pair = (method == MSED_RUNGE_KUTTA_4 ? RK4_12 : RK38_12) + which;
#define MSED_RKP(MODEL, PAIR) \
    case PAIR: rk_pair_kernel<MODEL, PAIR><<<grid, block, rk_pair_smem(PAIR), s>>>(p); break;
    if (model == MSED_MODEL_OMEXDIA_P) {
        switch (pair) {
            MSED_RKP(MSED_MODEL_OMEXDIA_P, RK4_12) MSED_RKP(MSED_MODEL_OMEXDIA_P, RK4_34)
            MSED_RKP(MSED_MODEL_OMEXDIA_P, RK38_12) MSED_RKP(MSED_MODEL_OMEXDIA_P, RK38_34)
        }
    } else {
        switch (pair) {
            MSED_RKP(MSED_MODEL_NONE, RK4_12) MSED_RKP(MSED_MODEL_NONE, RK4_34)
            MSED_RKP(MSED_MODEL_NONE, RK38_12) MSED_RKP(MSED_MODEL_NONE, RK38_34)
        }
    }
#undef MSED_RKP
    return cudaGetLastError();
}

cudaError_t tu_enable_rk_smem()
{
    cudaError_t e;
#define MSED_RKP_ATTR(MODEL, PAIR) \
    if ((e = cudaFuncSetAttribute(rk_pair_kernel<MODEL, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)rk_pair_smem(PAIR))) != cudaSuccess) return e;
    MSED_RKP_ATTR(MSED_MODEL_OMEXDIA_P, RK4_12) MSED_RKP_ATTR(MSED_MODEL_OMEXDIA_P, RK4_34)
    MSED_RKP_ATTR(MSED_MODEL_OMEXDIA_P, RK38_12) MSED_RKP_ATTR(MSED_MODEL_OMEXDIA_P, RK38_34)
    MSED_RKP_ATTR(MSED_MODEL_NONE, RK4_12) MSED_RKP_ATTR(MSED_MODEL_NONE, RK4_34)
    MSED_RKP_ATTR(MSED_MODEL_NONE, RK38_12) MSED_RKP_ATTR(MSED_MODEL_NONE, RK38_34)
#undef MSED_RKP_ATTR
    return cudaSuccess;
}

}  // namespace msed
