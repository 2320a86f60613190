// Build shim (ours): lets g++ parse reference translation units that are normally compiled as
// CUDA by clang. Force-included (-include) ahead of every reference source.
#pragma once
#include <algorithm>
#include <cassert>
#include <concepts>
#include <cstdint>
#include <cstring>
#include <memory_resource>
#include <utility>
inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
