// TEST INFRASTRUCTURE. Link stubs (ours) for the four symbols whose reference definitions need
// boost / spdlog / a CUDA kernel TU (sxt/base/error/stacktrace.cc, sxt/base/log/log_impl.cc,
// sxt/seqcommit/generator/gpu_generator.cc). None is reachable from the CPU MSM path.
#include <cstdint>
#include <cstdlib>
#include <string>
#include <string_view>

#include "sxt/base/container/span.h"
#include "sxt/curve21/type/element_p3.h"
#include "sxt/scalar25/operation/mul.h"
#include "sxt/scalar25/operation/muladd.h"
#include "sxt/scalar25/type/element.h"

namespace sxt::baser {
std::string stacktrace() noexcept { return "<no stacktrace in oracle build>"; }
} // namespace sxt::baser
namespace sxt::basl {
void info_impl(std::string_view) noexcept {}
void error_impl(std::string_view) noexcept {}
} // namespace sxt::basl
namespace sxt::sqcgn {
void gpu_get_generators(basct::span<c21t::element_p3>, uint64_t) noexcept { std::abort(); }
} // namespace sxt::sqcgn

// sxt/scalar25/operation/inner_product.cc also holds the GPU reduction (CUDA headers); the CPU
// overload is the plain multiply-accumulate loop of its lines 52-61, repeated here.
namespace sxt::s25o {
void inner_product(s25t::element& res, basct::cspan<s25t::element> lhs,
                   basct::cspan<s25t::element> rhs) noexcept {
  auto n = std::min(lhs.size(), rhs.size());
  s25o::mul(res, lhs[0], rhs[0]);
  for (size_t i = 1; i < n; ++i) {
    s25o::muladd(res, lhs[i], rhs[i], res);
  }
}
} // namespace sxt::s25o
