// msed_tu_pair_none.cu -- pair_kernel for the reaction-free model (see msed_tu_pair.inc), and the launcher
// that picks the model
#define MSED_TU_PAIR_MODEL MSED_MODEL_NONE
#define MSED_TU_PAIR_SUFFIX none
#include "msed_tu_pair.inc"

namespace msed {
cudaError_t tu_launch_pair_omexdia(bool adaptive, const KParams &p, cudaStream_t s);
cudaError_t tu_enable_pair_smem_omexdia();

cudaError_t tu_launch_pair(int model, bool adaptive, const KParams &p, cudaStream_t s)
{
    return model == MSED_MODEL_OMEXDIA_P ? tu_launch_pair_omexdia(adaptive, p, s) : tu_launch_pair_none(adaptive, p, s);
}

cudaError_t tu_enable_pair_smem()
{
    cudaError_t e = tu_enable_pair_smem_omexdia();
    return e != cudaSuccess ? e : tu_enable_pair_smem_none();
}
}  // namespace msed
