// Shadow header (ours): the reference's sxt/algorithm/iteration/for_each.h launches a CUDA kernel
// with <<< >>>, which g++ cannot parse. The CPU oracle never launches kernels, so the launcher
// traps if it is ever reached.
#pragma once
#include "sxt/algorithm/base/index_functor.h"
#include "sxt/base/device/stream.h"
#include "sxt/base/num/divide_up.h"
#include "sxt/base/type/raw_stream.h"
#include "sxt/execution/async/future.h"
#include "sxt/execution/device/synchronization.h"
#include "sxt/execution/kernel/kernel_dims.h"
namespace sxt::algi {
template <class F> void launch_for_each_kernel(bast::raw_stream_t, F, unsigned) noexcept {
  __builtin_trap();
}
template <class F> xena::future<> for_each(basdv::stream&&, F, unsigned) noexcept {
  __builtin_trap();
}
template <class F> xena::future<> for_each(F, unsigned) noexcept { __builtin_trap(); }
} // namespace sxt::algi
