// msed_tu_pair_omexdia.cu -- pair_kernel for the hzg_omexdia_p reaction model (see msed_tu_pair.inc)
#define MSED_TU_PAIR_MODEL MSED_MODEL_OMEXDIA_P
#define MSED_TU_PAIR_SUFFIX omexdia
#include "msed_tu_pair.inc"
