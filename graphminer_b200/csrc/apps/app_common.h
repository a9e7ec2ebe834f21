// Shared by the CLI drop-ins: argv handling identical to the reference mains.
#pragma once
#include "solvers.h"
using gm::Graph;
using gm::Pattern;
#ifndef GM_DEFAULT_NGPU
#define GM_DEFAULT_NGPU 1
#endif
