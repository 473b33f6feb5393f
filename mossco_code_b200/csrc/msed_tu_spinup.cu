// msed_tu_spinup.cu -- instantiations of spinup_kernel (msed_spinup.cuh) and their launcher.
#include "msed_launch.h"

namespace msed {
#include "msed_column.cuh"
#include "msed_spinup.cuh"

cudaError_t tu_launch_spinup(int model, const KParams &p, const SpinupArgs &a, cudaStream_t s)
{
    if (p.ncol <= 0) return cudaSuccess;
    const dim3 grid((p.ncol + SPINUP_WARPS - 1) / SPINUP_WARPS), block(SPINUP_BLOCK);
    if (p.K > 32) {   // two layers per lane
        if (model == MSED_MODEL_OMEXDIA_P) spinup_kernel<MSED_MODEL_OMEXDIA_P, 2><<<grid, block, 0, s>>>(p, a);
        else spinup_kernel<MSED_MODEL_NONE, 2><<<grid, block, 0, s>>>(p, a);
    } else {
        if (model == MSED_MODEL_OMEXDIA_P) spinup_kernel<MSED_MODEL_OMEXDIA_P, 1><<<grid, block, 0, s>>>(p, a);
        else spinup_kernel<MSED_MODEL_NONE, 1><<<grid, block, 0, s>>>(p, a);
    }
    return cudaGetLastError();
}

}  // namespace msed
