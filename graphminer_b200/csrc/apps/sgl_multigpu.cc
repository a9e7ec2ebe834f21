#define GM_DEFAULT_NGPU 8
#include "sgl_gpu_base.cc"
