#define GM_DEFAULT_NGPU 8
#include "tc_gpu_base.cc"
