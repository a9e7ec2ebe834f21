#define GM_DEFAULT_NGPU 8
#include "clique_gpu_base.cc"
