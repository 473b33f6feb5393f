// pair_kernel for the hzg_omexdia_p model, adaptive Euler (see msed_tu_pair.inc)
#define MSED_TU_PAIR_MODEL MSED_MODEL_OMEXDIA_P
#define MSED_TU_PAIR_ADAPTIVE true
#define MSED_TU_PAIR_SUFFIX omexdia_adaptive
#include "msed_tu_pair.inc"
