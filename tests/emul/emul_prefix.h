// TEST INFRASTRUCTURE — force-included when the product's CUDA sources are compiled as host C++ for
// the CPU emulation harness (see emul.cpp).
#pragma once
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#include <cstdint>
struct uint4 { uint32_t x, y, z, w; };
