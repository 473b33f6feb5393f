// pair_kernel for the reaction-free model, adaptive Euler (see msed_tu_pair.inc)
#define MSED_TU_PAIR_MODEL MSED_MODEL_NONE
#define MSED_TU_PAIR_ADAPTIVE true
#define MSED_TU_PAIR_SUFFIX none_adaptive
#include "msed_tu_pair.inc"
