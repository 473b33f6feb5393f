// msed_tu_rkquad.cu -- instantiations of rk_quad_kernel (msed_rkquad.cuh) and their launcher.
#include "msed_launch.h"

namespace msed {
#include "msed_column.cuh"
#include "msed_pair.cuh"   // sts64 and the shared-memory geometry
#include "msed_rkquad.cuh"

cudaError_t tu_launch_rk_quad(int model, int method, const KParams &p, cudaStream_t s)
{
    if (p.col_end <= p.col0) return cudaSuccess;
    if (p.K < RKQ_MIN_LAYERS) return cudaErrorInvalidValue;
    const dim3 grid((p.col_end - p.col0 + COL_BLOCK - 1) / COL_BLOCK), block(COL_BLOCK);
    const bool is38 = method == MSED_RUNGE_KUTTA_4_38;
    if (model == MSED_MODEL_OMEXDIA_P) {
        if (is38) rk_quad_kernel<MSED_MODEL_OMEXDIA_P, true><<<grid, block, RKQ_SMEM_BYTES, s>>>(p);
        else rk_quad_kernel<MSED_MODEL_OMEXDIA_P, false><<<grid, block, RKQ_SMEM_BYTES, s>>>(p);
    } else {
        if (is38) rk_quad_kernel<MSED_MODEL_NONE, true><<<grid, block, RKQ_SMEM_BYTES, s>>>(p);
        else rk_quad_kernel<MSED_MODEL_NONE, false><<<grid, block, RKQ_SMEM_BYTES, s>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t tu_enable_rk_quad_smem()
{
    cudaError_t e;
#define MSED_RKQ_ATTR(MODEL, IS38) \
    if ((e = cudaFuncSetAttribute(rk_quad_kernel<MODEL, IS38>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)RKQ_SMEM_BYTES)) != cudaSuccess) return e;
    MSED_RKQ_ATTR(MSED_MODEL_OMEXDIA_P, false) MSED_RKQ_ATTR(MSED_MODEL_OMEXDIA_P, true)
    MSED_RKQ_ATTR(MSED_MODEL_NONE, false) MSED_RKQ_ATTR(MSED_MODEL_NONE, true)
#undef MSED_RKQ_ATTR
    return cudaSuccess;
}

}  // namespace msed
