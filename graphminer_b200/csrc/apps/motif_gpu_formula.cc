#define GM_MOTIF_FORMULA 1
#include "motif_gpu_base.cc"
