// Shadow header (ours): the reference's sxt/multiexp/curve/multiexponentiation.h also pulls in the
// CUDA bucket-method headers, which g++ cannot parse. Callers on the CPU path (the inner-product
// cpu driver) only use compute_multiexponentiation, whose body (reference lines 128-142) is the
// ten-line composition of the reference's own cpu driver + Pippenger solver reproduced here.
#pragma once
#include "sxt/base/container/span.h"
#include "sxt/base/curve/element.h"
#include "sxt/memory/management/managed_array.h"
#include "sxt/multiexp/base/exponent_sequence.h"
#include "sxt/multiexp/curve/multiexponentiation_cpu_driver.h"
#include "sxt/multiexp/curve/pippenger_multiproduct_solver.h"
#include "sxt/multiexp/pippenger/multiexponentiation.h"

namespace sxt::mtxcrv {
template <bascrv::element Element>
memmg::managed_array<Element>
compute_multiexponentiation(basct::cspan<Element> generators,
                            basct::cspan<mtxb::exponent_sequence> exponents) noexcept {
  pippenger_multiproduct_solver<Element> solver;
  multiexponentiation_cpu_driver<Element> driver{&solver};
  return mtxpi::compute_multiexponentiation(
             driver,
             {static_cast<const void*>(generators.data()), generators.size(), sizeof(Element)},
             exponents)
      .value()
      .template as_array<Element>();
}
} // namespace sxt::mtxcrv
