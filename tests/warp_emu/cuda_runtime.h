// TEST HARNESS ONLY (tests/warp_emu): stands in for <cuda_runtime.h> when the kernel sources are
// compiled by g++ for the CPU warp emulator.  See cuda_shim.h.
#pragma once
#include "cuda_shim.h"
