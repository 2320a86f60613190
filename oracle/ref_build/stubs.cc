// TEST INFRASTRUCTURE. Link stubs (ours) for the four symbols whose reference definitions need
// boost / spdlog / a CUDA kernel TU (sxt/base/error/stacktrace.cc, sxt/base/log/log_impl.cc,
// sxt/seqcommit/generator/gpu_generator.cc). None is reachable from the CPU MSM path.
#include <cstdint>
#include <cstdlib>
#include <string>
#include <string_view>

#include "sxt/base/container/span.h"
#include "sxt/curve21/type/element_p3.h"

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
