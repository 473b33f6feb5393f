// pair_kernel for the reaction-free model, Euler (see msed_tu_pair.inc)
#define MSED_TU_PAIR_MODEL MSED_MODEL_NONE
#define MSED_TU_PAIR_ADAPTIVE false
#define MSED_TU_PAIR_SUFFIX none_euler
#include "msed_tu_pair.inc"

namespace msed {
cudaError_t tu_launch_pair_omexdia_adaptive(const KParams &p, cudaStream_t s);
cudaError_t tu_launch_pair_omexdia_euler(const KParams &p, cudaStream_t s);
cudaError_t tu_launch_pair_none_adaptive(const KParams &p, cudaStream_t s);
cudaError_t tu_enable_pair_smem_omexdia_adaptive();
cudaError_t tu_enable_pair_smem_omexdia_euler();
cudaError_t tu_enable_pair_smem_none_adaptive();

// the launcher that picks the translation unit
cudaError_t tu_launch_pair(int model, bool adaptive, const KParams &p, cudaStream_t s)
{
    if (model == MSED_MODEL_OMEXDIA_P) return adaptive ? tu_launch_pair_omexdia_adaptive(p, s) : tu_launch_pair_omexdia_euler(p, s);
    return adaptive ? tu_launch_pair_none_adaptive(p, s) : tu_launch_pair_none_euler(p, s);
}

cudaError_t tu_enable_pair_smem()
{
    cudaError_t e;
    if ((e = tu_enable_pair_smem_omexdia_adaptive()) != cudaSuccess) return e;
    if ((e = tu_enable_pair_smem_omexdia_euler()) != cudaSuccess) return e;
    if ((e = tu_enable_pair_smem_none_adaptive()) != cudaSuccess) return e;
    return tu_enable_pair_smem_none_euler();
}
}  // namespace msed
