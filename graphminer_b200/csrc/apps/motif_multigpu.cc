#define GM_DEFAULT_NGPU 8
#include "motif_gpu_base.cc"
