// Stand-in for <cuda_runtime.h> when the kernels are compiled for the host emulation (tests/emu): see cuda_emu.h.
#pragma once
#include "cuda_emu.h"
