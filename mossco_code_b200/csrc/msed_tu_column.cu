// msed_tu_column.cu -- instantiations of column_kernel (msed_column.cuh) and their launcher.
#include "msed_launch.h"

namespace msed {
#include "msed_column.cuh"

namespace {
inline int nblk(int n, int bs) { return (n + bs - 1) / bs; }

template <int MODEL, bool P3, bool SP>
cudaError_t launch_op(int op, const KParams &p, cudaStream_t s)
{
    const dim3 grid(nblk(p.col_end - p.col0, COL_BLOCK)), block(COL_BLOCK);
    switch (op) {
#define MSED_CASE(OPV) \
    case OPV: column_kernel<MODEL, OPV, P3, SP><<<grid, block, COLUMN_SMEM_BYTES, s>>>(p); break;
        MSED_CASE(OP_RHS)
        MSED_CASE(OP_EULER)
        MSED_CASE(OP_ADAPTIVE)
        MSED_CASE(OP_RK4_S1)
        MSED_CASE(OP_RK4_S2)
        MSED_CASE(OP_RK4_S3)
        MSED_CASE(OP_RK4_S4)
        MSED_CASE(OP_RK38_S1)
        MSED_CASE(OP_RK38_S2)
        MSED_CASE(OP_RK38_S3)
        MSED_CASE(OP_RK38_S4)
#undef MSED_CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
}  // namespace

cudaError_t tu_launch_column(int model, bool p3, bool sp, int op, const KParams &p, cudaStream_t s)
{
    switch (model) {
    case MSED_MODEL_OMEXDIA_P:
        if (p3) return launch_op<MSED_MODEL_OMEXDIA_P, true, true>(op, p, s);
        return sp ? launch_op<MSED_MODEL_OMEXDIA_P, false, true>(op, p, s)
                  : launch_op<MSED_MODEL_OMEXDIA_P, false, false>(op, p, s);
    case MSED_MODEL_NONE:
        if (p3) return launch_op<MSED_MODEL_NONE, true, true>(op, p, s);
        return sp ? launch_op<MSED_MODEL_NONE, false, true>(op, p, s)
                  : launch_op<MSED_MODEL_NONE, false, false>(op, p, s);
    case MSED_MODEL_TEST_SOLVER:
        if (op != OP_RHS && op != OP_EULER) return cudaErrorInvalidValue;
        return launch_op<MSED_MODEL_TEST_SOLVER, false, true>(op, p, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace msed
