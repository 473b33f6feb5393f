// pair_kernel for the hzg_omexdia_p model, Euler (see msed_tu_pair.inc)
#define MSED_TU_PAIR_MODEL MSED_MODEL_OMEXDIA_P
#define MSED_TU_PAIR_ADAPTIVE false
#define MSED_TU_PAIR_SUFFIX omexdia_euler
#include "msed_tu_pair.inc"
