// oracle/ref_compat_shim.h -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// Force-included (nvcc/g++ `-include`) in front of the UNMODIFIED reference
// sources houghvoting/src/hv_cuda_kernel.cu and hv_cuda.cpp so that they compile
// against torch 2.11 where they lie under /root/reference.
//
// The reference dispatches with `AT_DISPATCH_FLOATING_TYPES(points.type(), ...)`
// (hv_cuda_kernel.cu:142,157,285).  `Tensor::type()` returns
// at::DeprecatedTypeProperties, which modern AT_DISPATCH no longer accepts.
// We re-define the dispatch macro so that it accepts either a ScalarType or a
// DeprecatedTypeProperties.  No arithmetic of the reference is touched.
#pragma once
#include <torch/extension.h>

namespace cvb200_ref_shim {
inline c10::ScalarType to_scalar_type(c10::ScalarType t) { return t; }
inline c10::ScalarType to_scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace cvb200_ref_shim

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...)                                   \
  AT_DISPATCH_SWITCH(::cvb200_ref_shim::to_scalar_type(TYPE), NAME,                   \
                     AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
